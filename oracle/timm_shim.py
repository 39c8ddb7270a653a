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


# ---------------------------------------------------------------------------------------------------------
# MaxViT ('maxvit_tiny_rw_224'): module-based twin of oracle/maxvit_oracle.py under timm's parameter names.
# Written against different primitives than the functional oracle (nn.Conv2d / nn.BatchNorm2d / nn.LayerNorm
# modules, torchvision's WindowPartition / SwapAxes / SqueezeExcitation, F.scaled_dot_product_attention with the
# relative-position bias as additive mask) so the two restatements check each other.
# ---------------------------------------------------------------------------------------------------------
from torchvision.models import maxvit as tvmv
from torchvision.ops import SqueezeExcitation
import torch.nn.functional as F

from . import maxvit_oracle as MO


class _BNAct(nn.BatchNorm2d):
    def __init__(self, c, act):
        super().__init__(c, eps=MO.BN_EPS)
        self.act = nn.SiLU() if act else nn.Identity()      # parameter-free: no state-dict keys

    def forward(self, x):
        return self.act(super().forward(x))


class _Shortcut(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.pool = nn.AvgPool2d(2)
        self.expand = nn.Conv2d(cin, cout, 1, bias=False) if cin != cout else nn.Identity()

    def forward(self, x):
        return self.expand(self.pool(x))


class _MbConv(nn.Module):
    def __init__(self, cin, cout, stride, arch):
        super().__init__()
        mid = arch["expand"] * cin
        self.shortcut = _Shortcut(cin, cout) if stride == 2 else nn.Identity()
        self.pre_norm = _BNAct(cin, act=False)
        self.conv1_1x1 = nn.Conv2d(cin, mid, 1, bias=False)
        self.norm1 = _BNAct(mid, act=True)
        self.conv2_kxk = nn.Conv2d(mid, mid, 3, stride=stride, padding=1, groups=mid, bias=False)
        self.norm2 = _BNAct(mid, act=True)
        self.se = SqueezeExcitation(mid, mid // arch["se_div"], activation=nn.SiLU, scale_activation=nn.Sigmoid)
        self.conv3_1x1 = nn.Conv2d(mid, cout, 1, bias=False)

    def forward(self, x):
        y = self.conv1_1x1(self.pre_norm(x))
        y = self.norm2(self.conv2_kxk(self.norm1(y)))
        return self.conv3_1x1(self.se(y)) + self.shortcut(x)


class _RelPos(nn.Module):
    def __init__(self, win, heads):
        super().__init__()
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * win - 1) ** 2, heads))
        self.register_buffer("relative_position_index", tvmv._get_relative_position_index(win, win), persistent=False)
        self.n = win * win

    def get_bias(self):
        b = self.relative_position_bias_table[self.relative_position_index.view(-1)].view(self.n, self.n, -1)
        return b.permute(2, 0, 1).unsqueeze(0).contiguous()


class _Attn(nn.Module):
    def __init__(self, c, arch):
        super().__init__()
        self.dh = arch["dim_head"]
        self.heads = c // self.dh
        self.qkv = nn.Linear(c, 3 * c)
        self.rel_pos = _RelPos(arch["window"], self.heads)
        self.proj = nn.Linear(c, c)

    def forward(self, x):                                   # [B, G, N, C]
        B, G, N, C = x.shape
        qkv = self.qkv(x).view(B * G, N, self.heads, 3 * self.dh).transpose(1, 2)
        q, k, v = qkv.chunk(3, dim=3)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=self.rel_pos.get_bias())
        return self.proj(o.transpose(1, 2).reshape(B, G, N, C))


class _Mlp(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.fc1, self.act, self.fc2 = nn.Linear(c, 4 * c), nn.GELU(), nn.Linear(4 * c, c)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class _PartAttn(nn.Module):
    """NCHW in/out; partitioning by torchvision's WindowPartition (+ SwapAxes for the grid variant)."""

    def __init__(self, c, kind, arch):
        super().__init__()
        self.kind, self.win = kind, arch["window"]
        self.norm1, self.attn = nn.LayerNorm(c, eps=MO.LN_EPS), _Attn(c, arch)
        self.norm2, self.mlp = nn.LayerNorm(c, eps=MO.LN_EPS), _Mlp(c)
        self.part, self.depart, self.swap = tvmv.WindowPartition(), tvmv.WindowDepartition(), tvmv.SwapAxes(-2, -3)

    def forward(self, x):
        H, W = x.shape[-2:]
        p = self.win if self.kind == "block" else H // self.win
        x = self.part(x, p)
        if self.kind == "grid":
            x = self.swap(x)
        x = x + self.attn(self.norm1(x))
        x = x + self.mlp(self.norm2(x))
        if self.kind == "grid":
            x = self.swap(x)
        return self.depart(x, p, H // p, W // p)


class _MaxVitBlock(nn.Module):
    def __init__(self, cin, cout, stride, arch):
        super().__init__()
        self.conv = _MbConv(cin, cout, stride, arch)
        self.attn_block = _PartAttn(cout, "block", arch)
        self.attn_grid = _PartAttn(cout, "grid", arch)

    def forward(self, x):
        return self.attn_grid(self.attn_block(self.conv(x)))


class _MaxVitStage(nn.Module):
    def __init__(self, cin, cout, depth, arch):
        super().__init__()
        self.blocks = nn.Sequential(*[_MaxVitBlock(cin if j == 0 else cout, cout, 2 if j == 0 else 1, arch)
                                      for j in range(depth)])

    def forward(self, x):
        return self.blocks(x)


class _MaxVitStem(nn.Module):
    def __init__(self, w):
        super().__init__()
        self.conv1 = nn.Conv2d(3, w[0], 3, stride=2, padding=1, bias=False)
        self.norm1 = _BNAct(w[0], act=True)
        self.conv2 = nn.Conv2d(w[0], w[1], 3, stride=1, padding=1, bias=False)

    def forward(self, x):
        return self.conv2(self.norm1(self.conv1(x)))


class _PoolFlatten(nn.Module):
    def forward(self, x):                                   # timm SelectAdaptivePool2d('avg', flatten=True)
        return x.mean(dim=(2, 3))


class _MaxVitHead(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.global_pool = _PoolFlatten()
        self.fc = nn.Linear(c, 1000)
        self.in_features = c

    def forward(self, x):
        return self.fc(self.global_pool(x))


class ShimMaxViT(nn.Module):
    def __init__(self, arch):
        super().__init__()
        dims = arch["embed_dim"]
        self.stem = _MaxVitStem(arch["stem_width"])
        cins = (arch["stem_width"][1],) + tuple(dims[:-1])
        self.stages = nn.Sequential(*[_MaxVitStage(cins[i], dims[i], arch["depths"][i], arch) for i in range(4)])
        self.norm = TVLayerNorm2d(dims[-1], eps=MO.LN_EPS)
        self.head = _MaxVitHead(dims[-1])

    def forward(self, x):
        return self.head(self.norm(self.stages(self.stem(x))))


def create_model(model_kind, pretrained=False, **kw):
    if pretrained:
        raise RuntimeError("timm shim: pretrained weights need the network")
    if "maxvit" in model_kind.lower():
        return ShimMaxViT(MO.arch_of(model_kind))
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
    if any("maxvit" in k for k in sd_timm):                 # the MaxViT twin already uses timm's key names
        sd = {k: torch.as_tensor(v) for k, v in sd_timm.items()}
        missing = [k for k in model.state_dict() if k not in sd]
        assert all(".head.fc." in k for k in missing), missing
        assert not [k for k in sd if k not in model.state_dict()]
        model.load_state_dict(sd, strict=False)
        return model
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
