"""Model classes with the reference's names, constructor config keys, forward keywords and state-dict keys
(`btsbot/architectures.py:25-372`), running on hand-written sm_100a kernels.

The ``nn.Module`` tree below exists to *hold parameters under timm's / the reference's key names* so that
``state_dict()`` / ``load_state_dict(strict=True)`` / ``.to()`` / ``.parameters()`` / ``DataParallel(...).module``
behave exactly like the reference's models.  None of the sub-modules' own ``forward`` methods is ever used:
``forward`` hands the whole batch to :class:`btsbot_b200._engine.Scorer`, i.e. to ``libbtsbot_b200.so``.
There is no CPU / eager fallback -- calling a model on CPU tensors raises.

Extra (opt-in) config keys: ``precision`` = ``"fp32"`` (default, reference numerics: logits within 1e-4; GEMMs on the
tensor cores through the 3xTF32 split) or ``"bf16"`` (tcgen05 bf16, logits within 2e-2); ``infer_cuda_graph`` = true
replays the eval-mode forward as one CUDA graph per input shape (small batches are otherwise paced by the host).
"""
from __future__ import annotations

import json
import os.path as path
import re
import warnings

import torch
import torch.nn as nn

from . import _engine
from .synth import convnext_arch


def get_model_image_size(model_kind: str) -> int:
    """Image size encoded in a MaxViT model name, else 224 (architectures.py:10-22)."""
    if "maxvit" in model_kind.lower():
        m = re.search(r"_(\d+)\.", model_kind)
        if m:
            return int(m.group(1))
    return 224


# ---------------------------------------------------------------------------------------------------------
# parameter containers mirroring timm's ConvNeXt module tree
# ---------------------------------------------------------------------------------------------------------
class _KernelOnly(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError("btsbot_b200: sub-modules are parameter containers; call the model, which runs the "
                           "fused sm_100a kernels")


class LayerNorm2d(_KernelOnly):
    def __init__(self, c: int, eps: float = 1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.normalized_shape = (c,)
        self.eps = eps


class _ConvParams(_KernelOnly):
    """Conv2d weight/bias holder (same parameter names and shapes as ``nn.Conv2d``)."""

    def __init__(self, cin, cout, k, groups=1):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin // groups, k, k))
        self.bias = nn.Parameter(torch.zeros(cout))
        self.in_channels, self.out_channels, self.kernel_size, self.groups = cin, cout, (k, k), groups
        nn.init.trunc_normal_(self.weight, std=0.02)       # timm ConvNeXt _init_weights


class _Mlp(_KernelOnly):
    def __init__(self, c):
        super().__init__()
        self.fc1 = _ConvParams(c, 4 * c, 1)
        self.fc2 = _ConvParams(4 * c, c, 1)


class _Block(_KernelOnly):
    def __init__(self, c, ls_init_value=1e-6):
        super().__init__()
        self.conv_dw = _ConvParams(c, c, 7, groups=c)
        self.norm = LayerNorm2d(c)
        self.mlp = _Mlp(c)
        self.gamma = nn.Parameter(ls_init_value * torch.ones(c))


class _Stage(_KernelOnly):
    def __init__(self, cin, c, depth, first):
        super().__init__()
        if first:
            self.downsample = nn.Identity()
        else:
            self.downsample = nn.Sequential(LayerNorm2d(cin), _ConvParams(cin, c, 2))
        self.blocks = nn.Sequential(*[_Block(c) for _ in range(depth)])


class _TimmHead(_KernelOnly):
    """Attribute surface of timm's ``NormMlpClassifierHead`` that the reference reads
    (architectures.py:109-113,134-143)."""

    def __init__(self, c):
        super().__init__()
        self.global_pool = nn.AdaptiveAvgPool2d(1)
        self.norm = LayerNorm2d(c)
        self.flatten = nn.Flatten(1)
        self.in_features = c


class ConvNeXtTrunk(nn.Module):
    """Stand-in for ``timm.create_model('convnext_nano|pico...')``: same parameter tree, no arithmetic."""

    def __init__(self, model_kind: str, pretrained: bool = False):
        super().__init__()
        arch = convnext_arch(model_kind)
        self.model_kind, self.arch = model_kind, arch
        if pretrained:
            warnings.warn("btsbot_b200: pretrained timm weights need the network; the trunk is random-initialised "
                          "-- load a checkpoint with load_state_dict()", stacklevel=3)
        dims, depths = arch["dims"], arch["depths"]
        self.stem = nn.Sequential(_ConvParams(3, dims[0], 4), LayerNorm2d(dims[0]))
        self.stages = nn.Sequential(*[
            _Stage(dims[i - 1] if i else dims[0], dims[i], depths[i], first=(i == 0)) for i in range(4)])
        self.norm_pre = nn.Identity()
        self.head = _TimmHead(dims[-1])
        self.num_features = dims[-1]

    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError("btsbot_b200: the trunk runs inside the model's fused forward")


# ---------------------------------------------------------------------------------------------------------
# shared forward machinery
# ---------------------------------------------------------------------------------------------------------
class _B200Model(nn.Module):
    """Caches a packed :class:`_engine.Scorer` and rebuilds it when parameters change or move.

    Invalidation is explicit where autograd's version counters cannot see the change: ``FusedAdamW.step`` and CUDA-graph
    replays write parameters through raw device pointers, so they bump ``_engine.param_generation`` (part of the cache
    key); ``train()`` / ``load_state_dict()`` / ``_apply()`` (``.to()``, ``.cuda()``, ``.half()`` ...) drop the cache."""

    def _init_runtime(self, config: dict):
        self._config = dict(config)
        self._precision = config.get("precision", "fp32")
        self._scorer = None
        self._scorer_key = None

    def _state_key(self):
        ver, n, dev = 0, 0, None
        for t in self.parameters():
            ver += t._version
            n += 1
            dev = t.device
        for t in self.buffers():
            ver += t._version
        return (ver, str(dev), self._precision, n, _engine.param_generation())

    def _drop_scorer(self):
        if getattr(self, "_scorer", None) is not None:
            self._scorer = None
            self._scorer_key = None
        if getattr(self, "_frozen_scorer", None) is not None:
            self._frozen_scorer = None

    def train(self, mode: bool = True):
        self._drop_scorer()
        return super().train(mode)

    def load_state_dict(self, *args, **kwargs):
        self._drop_scorer()
        return super().load_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._drop_scorer()
        return super()._apply(fn, *args, **kwargs)

    def set_precision(self, precision: str):
        """``"fp32"`` or ``"bf16"``; takes effect on the next forward."""
        if precision not in ("fp32", "bf16"):
            raise ValueError(precision)
        self._precision = precision
        return self

    def scorer(self) -> _engine.Scorer:
        key = self._state_key()
        if self._scorer is None or key != self._scorer_key:
            self._scorer = _engine.Scorer(self._config, self.state_dict(), self._precision)
            self._scorer_key = key
        return self._scorer

    def _frozen_image_scorer(self) -> _engine.Scorer:
        """Scorer whose ``features()`` the training path uses while the image trunk is frozen (train.py:224-231).  Keyed
        on the frozen parameters only -- the trainable ones (and BatchNorm's ``num_batches_tracked``) change every step
        and would otherwise force a re-pack of the whole trunk per step (and inside a CUDA-graph capture)."""
        ver, dev = 0, None
        for t in self.parameters():
            if not t.requires_grad:
                ver += t._version
                dev = t.device
        key = (ver, str(dev), self._precision)
        if getattr(self, "_frozen_scorer", None) is None or key != self._frozen_key:
            self._frozen_scorer = _engine.Scorer(self._config, self.state_dict(), self._precision)
            self._frozen_key = key
        return self._frozen_scorer

    def _run(self, image_input=None, metadata_input=None):
        # kernels are enqueued on the CURRENT device's current stream: make the inputs' device current for the call
        # (a process that drives several GPUs may call a model whose tensors live on a non-current device)
        t = image_input if image_input is not None else metadata_input
        if t is not None and t.is_cuda and t.device.index != torch.cuda.current_device():
            with torch.cuda.device(t.device):
                return self._run_on_device(image_input, metadata_input)
        return self._run_on_device(image_input, metadata_input)

    def _run_on_device(self, image_input=None, metadata_input=None):
        if self.training and torch.is_grad_enabled():
            from . import _autograd
            return _autograd.training_forward(self, image_input, metadata_input)
        if self._config.get("infer_cuda_graph", False):
            # opt-in: replay the whole forward as one CUDA graph per input shape (small-batch scoring is host-paced)
            return self.scorer().graphed(image_input=image_input, metadata_input=metadata_input)
        return self.scorer()(image_input=image_input, metadata_input=metadata_input)


def _metadata_branch(n, cfg, act):
    # architectures.py:146-153 (GELU) / :205-212, :282-289 (ReLU)
    return [nn.BatchNorm1d(n), nn.Linear(n, cfg["meta_fc1_neurons"]), act(), nn.Dropout(cfg["meta_dropout"]),
            nn.Linear(cfg["meta_fc1_neurons"], cfg["meta_fc2_neurons"]), act()]


def _combined_head(nin, cfg, act):
    # architectures.py:157-164 / :357-365
    return nn.Sequential(nn.Linear(nin, cfg["comb_fc1_neurons"]), act(),
                         nn.Linear(cfg["comb_fc1_neurons"], cfg["comb_fc2_neurons"]), act(),
                         nn.Dropout(cfg["comb_dropout"]), nn.Linear(cfg["comb_fc2_neurons"], 1))


# ---------------------------------------------------------------------------------------------------------
# reference model classes
# ---------------------------------------------------------------------------------------------------------
class ConvNeXt(_B200Model):
    """Image-only ConvNeXt (architectures.py:104-122): trunk -> pool -> LN2d -> flatten -> 3-layer head."""

    def __init__(self, config):
        super().__init__()
        model_kind = config.get("model_kind", "convnext_nano.d1h_in1k")
        self.convnext = ConvNeXtTrunk(model_kind, pretrained=config.get("pretrained", True))
        h = self.convnext.head
        self.convnext.head = nn.Sequential(
            h.global_pool, h.norm, h.flatten,
            nn.Linear(h.in_features, config["fc1_neurons"]), nn.GELU(),
            nn.Linear(config["fc1_neurons"], config["fc2_neurons"]), nn.GELU(),
            nn.Dropout(config["dropout"]), nn.Linear(config["fc2_neurons"], 1))
        self._init_runtime(dict(config, model_name="ConvNeXt"))

    def forward(self, input_data: torch.Tensor) -> torch.Tensor:
        return self._run(image_input=input_data)


class mm_ConvNeXt(_B200Model):
    """Multimodal ConvNeXt (architectures.py:125-171)."""

    def __init__(self, config):
        super().__init__()
        model_kind = config.get("model_kind", "convnext_nano.d1h_in1k")
        n_meta = len(config.get("metadata_cols", []))
        self.convnext_backbone = ConvNeXtTrunk(model_kind, pretrained=config.get("pretrained", True))
        self.convnext_feature_dim = self.convnext_backbone.head.in_features
        h = self.convnext_backbone.head
        if "LS" in config["train_data_version"]:
            self.convnext_backbone.head = nn.Sequential(h.global_pool, h.norm, h.flatten)
        else:
            self.convnext_backbone.head = h.flatten
        self.metadata_branch = nn.Sequential(*_metadata_branch(n_meta, config, nn.GELU))
        self.combined_head = _combined_head(self.convnext_feature_dim + config["meta_fc2_neurons"], config, nn.GELU)
        self._init_runtime(dict(config, model_name="mm_ConvNeXt"))

    def forward(self, image_input: torch.Tensor, metadata_input: torch.Tensor) -> torch.Tensor:
        return self._run(image_input=image_input, metadata_input=metadata_input)


class um_nn(_B200Model):
    """Metadata-only MLP (architectures.py:277-293); runs in the fused metadata/head kernel."""

    def __init__(self, config):
        super().__init__()
        n_meta = len(config.get("metadata_cols", []))
        self.network = nn.Sequential(*_metadata_branch(n_meta, config, nn.ReLU),
                                     nn.Linear(config["meta_fc2_neurons"], 1))
        self._init_runtime(dict(config, model_name="um_nn"))

    def forward(self, input_data: torch.Tensor) -> torch.Tensor:
        return self._run(metadata_input=input_data)


class frozen_fusion(_B200Model):
    """Late fusion of a trained image branch and metadata branch (architectures.py:296-372) -- what the
    published HF "-metadata" checkpoints are (`to_HF.py:143`)."""

    @staticmethod
    def remove_branch_head(model, model_name):
        if model_name == "um_nn":
            model.network = nn.Sequential(*list(model.network.children())[:-2])
            emb_dim = model.network[-1].out_features
        elif model_name == "ConvNeXt":
            model.convnext.head = nn.Sequential(*list(model.convnext.head.children())[0:3])
            emb_dim = model.convnext.head[1].normalized_shape[0]
        elif model_name == "MaxViT":
            emb_dim = model.maxvit.head[1].in_features
            model.maxvit.head = nn.Sequential(*list(model.maxvit.head.children())[0:1])
        elif model_name == "um_cnn":
            emb_dim = model.head[0].in_features
            model.head = nn.Identity()
        else:
            raise ValueError(f"Model {model_name} not supported")
        return model, emb_dim

    @staticmethod
    def load_BTSbot_model(model_dir, train_config=None, skip_load_state=False):
        if train_config is None:
            with open(path.join(model_dir, "report.json"), "r") as f:
                train_config = json.load(f)["train_config"]
        try:
            model_type = globals()[train_config["model_name"]]
        except KeyError:
            print(f"Could not find model of name {train_config['model_name']}")
            exit(0)
        model = model_type(train_config)
        if not skip_load_state:
            model.load_state_dict(torch.load(path.join(model_dir, "best_model.pth")))
        return frozen_fusion.remove_branch_head(model, train_config["model_name"])

    def __init__(self, config):
        super().__init__()
        skip = config.get("skip_load_state", False)
        self.image_branch, img_dim = frozen_fusion.load_BTSbot_model(
            config["image_model_dir"], train_config=config.get("image_model_config", None), skip_load_state=skip)
        self.meta_branch, meta_dim = frozen_fusion.load_BTSbot_model(
            config["meta_model_dir"], train_config=config.get("meta_model_config", None), skip_load_state=skip)
        self.combined_head = _combined_head(img_dim + meta_dim, config, nn.ReLU)
        cfg = dict(config, model_name="frozen_fusion")
        cfg["image_model_config"] = dict(self.image_branch._config)
        cfg["meta_model_config"] = dict(self.meta_branch._config)
        self._init_runtime(cfg)

    def forward(self, image_input: torch.Tensor, metadata_input: torch.Tensor) -> torch.Tensor:
        return self._run(image_input=image_input, metadata_input=metadata_input)


# ---------------------------------------------------------------------------------------------------------
# legacy models (architectures.py:174-274): outside the north-star hot path.  SURVEY.md section 8 row a8 keeps them as
# plain PyTorch pass-through classes so that the package's API surface (btsbot/__init__.py:16-25) and old checkpoints'
# state-dict keys stay complete; they run on PyTorch's own ops (eager), not on this package's kernels.
# ---------------------------------------------------------------------------------------------------------
def _legacy_conv_stack(cfg) -> nn.Sequential:
    """Two [conv, ReLU, conv, ReLU, max-pool, Dropout2d] groups + Flatten: 5x5 'same' convs 3 -> c1 -> c1 -> pool 2 ->
    c2 -> c2 -> pool 4 (Sequential indices 0-12 are the checkpoint key contract)."""
    k, c1, c2 = cfg["conv_kernel"], cfg["conv1_channels"], cfg["conv2_channels"]
    layers = []
    for cin, cout, pool, drop in ((3, c1, 2, cfg["conv_dropout1"]), (c1, c2, 4, cfg["conv_dropout2"])):
        layers += [nn.Conv2d(cin, cout, kernel_size=k, padding="same"), nn.ReLU(),
                   nn.Conv2d(cout, cout, kernel_size=k, padding="same"), nn.ReLU(),
                   nn.MaxPool2d(kernel_size=pool, stride=pool), nn.Dropout2d(drop)]
    return nn.Sequential(*layers, nn.Flatten())


def _legacy_feature_dim(cfg) -> int:
    return cfg["conv2_channels"] * (cfg.get("image_size", 63) // 8) ** 2


class mm_cnn(nn.Module):
    """Legacy multimodal CNN (architectures.py:174-229), PyTorch pass-through."""

    def __init__(self, config):
        super().__init__()
        n_meta = len(config.get("metadata_cols", []))
        self.conv_layers = _legacy_conv_stack(config)
        self.conv_feature_dim = _legacy_feature_dim(config)
        self.metadata_branch = nn.Sequential(*_metadata_branch(n_meta, config, nn.ReLU))
        self.combined_head = _combined_head(self.conv_feature_dim + config["meta_fc2_neurons"], config, nn.ReLU)
        self._config = dict(config, model_name="mm_cnn")

    def forward(self, image_input: torch.Tensor, metadata_input: torch.Tensor) -> torch.Tensor:
        both = torch.cat((self.conv_layers(image_input), self.metadata_branch(metadata_input)), dim=1)
        return self.combined_head(both)


class um_cnn(nn.Module):
    """Legacy image-only CNN (architectures.py:232-274), PyTorch pass-through."""

    def __init__(self, config):
        super().__init__()
        self.conv_layers = _legacy_conv_stack(config)
        self.head = nn.Sequential(
            nn.Linear(_legacy_feature_dim(config), config["fc1_neurons"]), nn.ReLU(),
            nn.Linear(config["fc1_neurons"], config["fc2_neurons"]), nn.ReLU(),
            nn.Dropout(config["dropout"]), nn.Linear(config["fc2_neurons"], 1))
        self._config = dict(config, model_name="um_cnn")

    def forward(self, input_data: torch.Tensor) -> torch.Tensor:
        return self.head(self.conv_layers(input_data))


# ---------------------------------------------------------------------------------------------------------
# parameter containers mirroring timm's MaxxVit module tree; forward never used
# ---------------------------------------------------------------------------------------------------------
class _RelPosBias(_KernelOnly):
    def __init__(self, win, heads):
        super().__init__()
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * win - 1) ** 2, heads))
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)


class _AttentionCl(_KernelOnly):
    def __init__(self, c, arch):
        super().__init__()
        self.qkv = nn.Linear(c, 3 * c)
        self.rel_pos = _RelPosBias(arch["window"], c // arch["dim_head"])
        self.proj = nn.Linear(c, c)


class _TokenMlp(_KernelOnly):
    def __init__(self, c):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(c, 4 * c), nn.Linear(4 * c, c)


class _PartitionAttention(_KernelOnly):
    def __init__(self, c, arch):
        super().__init__()
        self.norm1, self.attn = nn.LayerNorm(c, eps=1e-6), _AttentionCl(c, arch)
        self.norm2, self.mlp = nn.LayerNorm(c, eps=1e-6), _TokenMlp(c)


class _SE(_KernelOnly):
    def __init__(self, c, rd):
        super().__init__()
        self.fc1, self.fc2 = nn.Conv2d(c, rd, 1), nn.Conv2d(rd, c, 1)


class _ShortcutParams(_KernelOnly):
    def __init__(self, cin, cout):
        super().__init__()
        self.expand = nn.Conv2d(cin, cout, 1, bias=False) if cin != cout else nn.Identity()


class _MbConv(_KernelOnly):
    def __init__(self, cin, cout, stride, arch):
        super().__init__()
        mid = arch["expand"] * cin
        self.shortcut = _ShortcutParams(cin, cout) if stride == 2 else nn.Identity()
        self.pre_norm = nn.BatchNorm2d(cin, eps=1e-5)
        self.conv1_1x1 = nn.Conv2d(cin, mid, 1, bias=False)
        self.norm1 = nn.BatchNorm2d(mid, eps=1e-5)
        self.conv2_kxk = nn.Conv2d(mid, mid, 3, stride=stride, padding=1, groups=mid, bias=False)
        self.norm2 = nn.BatchNorm2d(mid, eps=1e-5)
        self.se = _SE(mid, mid // arch["se_div"])
        self.conv3_1x1 = nn.Conv2d(mid, cout, 1, bias=False)


class _MaxVitBlock(_KernelOnly):
    def __init__(self, cin, cout, stride, arch):
        super().__init__()
        self.conv = _MbConv(cin, cout, stride, arch)
        self.attn_block = _PartitionAttention(cout, arch)
        self.attn_grid = _PartitionAttention(cout, arch)


class _MaxVitStage(_KernelOnly):
    def __init__(self, cin, cout, depth, arch):
        super().__init__()
        self.blocks = nn.Sequential(*[_MaxVitBlock(cin if j == 0 else cout, cout, 2 if j == 0 else 1, arch)
                                      for j in range(depth)])


class _MaxVitStem(_KernelOnly):
    def __init__(self, widths):
        super().__init__()
        self.conv1 = nn.Conv2d(3, widths[0], 3, stride=2, padding=1, bias=False)
        self.norm1 = nn.BatchNorm2d(widths[0], eps=1e-5)
        self.conv2 = nn.Conv2d(widths[0], widths[1], 3, stride=1, padding=1, bias=False)


class _GlobalPool(_KernelOnly):
    """Stands in for timm's ``SelectAdaptivePool2d('avg', flatten=True)`` (parameter-free)."""


class _MaxVitHead(_KernelOnly):
    """Attribute surface of timm's ``ClassifierHead`` that the reference reads (architectures.py:33-34,64-65)."""

    def __init__(self, c):
        super().__init__()
        self.global_pool = _GlobalPool()
        self.in_features = c


class MaxVitTrunk(nn.Module):
    """Stand-in for ``timm.create_model('maxvit_tiny_rw_224...')``: same parameter tree, no arithmetic."""

    def __init__(self, model_kind: str, pretrained: bool = False):
        super().__init__()
        from .synth import maxvit_arch
        arch = maxvit_arch(model_kind)
        self.model_kind, self.arch = model_kind, arch
        if pretrained:
            warnings.warn("btsbot_b200: pretrained timm weights need the network; the trunk is random-initialised "
                          "-- load a checkpoint with load_state_dict()", stacklevel=3)
        dims = arch["embed_dim"]
        self.stem = _MaxVitStem(arch["stem_width"])
        cins = (arch["stem_width"][1],) + tuple(dims[:-1])
        self.stages = nn.Sequential(*[_MaxVitStage(cins[i], dims[i], arch["depths"][i], arch) for i in range(4)])
        self.norm = LayerNorm2d(dims[-1])
        self.head = _MaxVitHead(dims[-1])
        self.num_features = dims[-1]

    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError("btsbot_b200: the trunk runs inside the model's fused forward")


class MaxViT(_B200Model):
    """Image-only MaxViT (architectures.py:25-51): bilinear resize to the model's image size, trunk, head =
    pool -> Linear -> GELU -> Linear -> GELU -> Dropout -> Linear."""

    def __init__(self, config):
        super().__init__()
        model_kind = config.get("model_kind", "maxvit_tiny_rw_224.sw_in1k")
        self.image_size = get_model_image_size(model_kind)
        self.maxvit = MaxVitTrunk(model_kind, pretrained=config.get("pretrained", True))
        h = self.maxvit.head
        self.maxvit.head = nn.Sequential(
            h.global_pool,
            nn.Linear(h.in_features, config["fc1_neurons"]), nn.GELU(),
            nn.Linear(config["fc1_neurons"], config["fc2_neurons"]), nn.GELU(),
            nn.Dropout(config["dropout"]), nn.Linear(config["fc2_neurons"], 1))
        self._init_runtime(dict(config, model_name="MaxViT"))

    def forward(self, input_data: torch.Tensor) -> torch.Tensor:
        return self._run(image_input=input_data)


class mm_MaxViT(_B200Model):
    """Multimodal MaxViT (architectures.py:54-101)."""

    def __init__(self, config):
        super().__init__()
        model_kind = config.get("model_kind", "maxvit_tiny_rw_224.sw_in1k")
        self.image_size = get_model_image_size(model_kind)
        n_meta = len(config.get("metadata_cols", []))
        self.maxvit_backbone = MaxVitTrunk(model_kind, pretrained=config.get("pretrained", True))
        self.maxvit_feature_dim = self.maxvit_backbone.head.in_features
        self.maxvit_backbone.head = self.maxvit_backbone.head.global_pool
        self.metadata_branch = nn.Sequential(*_metadata_branch(n_meta, config, nn.GELU))
        self.combined_head = _combined_head(self.maxvit_feature_dim + config["meta_fc2_neurons"], config, nn.GELU)
        self._init_runtime(dict(config, model_name="mm_MaxViT"))

    def forward(self, image_input: torch.Tensor, metadata_input: torch.Tensor) -> torch.Tensor:
        return self._run(image_input=image_input, metadata_input=metadata_input)
