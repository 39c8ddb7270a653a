"""Build-container-only helper (TEST INFRASTRUCTURE): lets `/root/reference/btsbot/architectures.py`
execute *verbatim* although ``timm`` is not installable offline.

``install()`` puts a module named ``timm`` in ``sys.modules`` whose ``create_model`` returns a trunk with
the attribute surface the reference touches (`architectures.py:33-34,64-65,109-113,134-143`):
``.head.{global_pool,norm,flatten,in_features,fc}`` and ``forward = head(features(x))``.  The trunk
arithmetic is torchvision's independently written ``ConvNeXt`` (same math as timm's nano/pico with
``conv_mlp=True``), so goldens made through this shim cross-check the oracle's restated trunk against
a second implementation.  ``load_timm_keys`` maps a timm-keyed state dict onto the torchvision modules.
"""
import re
import sys
import types

import torch
import torch.nn as nn
from torchvision.models.convnext import CNBlockConfig, ConvNeXt as TVConvNeXt, LayerNorm2d as TVLayerNorm2d
from functools import partial

from .convnext_oracle import arch_of


class _Head(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.global_pool = nn.AdaptiveAvgPool2d(1)
        self.norm = TVLayerNorm2d(c, eps=1e-6)
        self.flatten = nn.Flatten(1)
        self.fc = nn.Linear(c, 1000)
        self.in_features = c

    def forward(self, x):
        return self.fc(self.flatten(self.norm(self.global_pool(x))))


class ShimConvNeXt(nn.Module):
    def __init__(self, arch):
        super().__init__()
        dims, depths = arch["dims"], arch["depths"]
        setting = [CNBlockConfig(dims[i], dims[i + 1] if i < 3 else None, depths[i]) for i in range(4)]
        tv = TVConvNeXt(setting, layer_scale=1e-6, norm_layer=partial(TVLayerNorm2d, eps=1e-6))
        self.features = tv.features
        self.head = _Head(dims[-1])

    def forward(self, x):
        return self.head(self.features(x))


def create_model(model_kind, pretrained=False, **kw):
    if pretrained:
        raise RuntimeError("timm shim: pretrained weights need the network")
    return ShimConvNeXt(arch_of(model_kind))


def install():
    m = types.ModuleType("timm")
    m.create_model = create_model
    m.__version__ = "shim"
    sys.modules["timm"] = m
    return m


def timm_to_tv_key(k: str) -> str:
    """``...stem.0.weight`` -> ``...features.0.0.weight`` etc. (trunk keys only; others returned unchanged)."""
    k = re.sub(r"\bstem\.(\d)\.", r"features.0.\1.", k)
    k = re.sub(r"\bstages\.(\d)\.downsample\.(\d)\.", lambda m: f"features.{2 * int(m.group(1))}.{m.group(2)}.", k)

    def blk(m):
        i, j, rest = int(m.group(1)), m.group(2), m.group(3)
        table = {"conv_dw": "block.0", "norm": "block.2", "mlp.fc1": "block.3", "mlp.fc2": "block.5"}
        if rest == "gamma":
            return f"features.{2 * i + 1}.{j}.layer_scale"
        for a, b in table.items():
            if rest.startswith(a + "."):
                return f"features.{2 * i + 1}.{j}.{b}.{rest[len(a) + 1:]}"
        raise KeyError(m.group(0))
    return re.sub(r"\bstages\.(\d)\.blocks\.(\d+)\.([\w.]+)$", blk, k)


def load_timm_keys(model: nn.Module, sd_timm: dict):
    """Load a timm/reference-keyed state dict into a reference model built on the shim (strict)."""
    out = {}
    for k, v in sd_timm.items():
        nk = timm_to_tv_key(k)
        v = torch.as_tensor(v)
        if nk.endswith("layer_scale"):
            v = v.reshape(-1, 1, 1)
        elif re.search(r"block\.[35]\.weight$", nk):
            v = v.reshape(v.shape[0], v.shape[1])      # Conv2d 1x1 [O,I,1,1] -> Linear [O,I]
        out[nk] = v
    target = model.state_dict()
    # the shim trunk keeps timm's discarded classifier (head.fc) only where the reference keeps the head object
    missing = [k for k in target if k not in out]
    extra = [k for k in out if k not in target]
    assert not extra, extra
    assert all(".fc." in k for k in missing), missing
    model.load_state_dict(out, strict=False)
    return model
