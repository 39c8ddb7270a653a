"""CPU fp32 restatement of the reference's MaxViT path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows
* `btsbot/architectures.py:25-51`   ``MaxViT``    (bilinear 63->224, trunk, head = pool -> Linear/GELU x2 -> Dropout -> Linear)
* `btsbot/architectures.py:54-101`  ``mm_MaxViT`` (bilinear 63->224, trunk, head = global_pool; BN1d-MLP; cat; fusion head)
* timm ``MaxxVit`` kind ``maxvit_tiny_rw_224`` (file ``maxxvit.py``; ``timm>=0.9.0`` per `pyproject.toml:43`, un-vendored and not
  installable offline) restated from SURVEY.md Appendix A.2.  **Parity vs timm itself is unpinned**: the reference holds
  no tests/golden vectors for it.  What is checked instead: analytic parameter count 29.06 M (timm's published figure),
  window/grid partition and the relative-position index against torchvision's independently written MaxViT helpers
  (tests/test_oracle_golden.py), and the reference's own wrapper code executed verbatim on a module-based twin of this
  restatement (oracle/timm_shim.py ``ShimMaxViT``).

Functional and state-dict driven (timm key names), NCHW in/out like timm.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .convnext_oracle import bn1d_eval, head3, lin, metadata_branch  # reference-owned glue is shared

LN_EPS = 1e-6      # transformer_cfg.norm_eps (LayerNorm / LayerNorm2d)
BN_EPS = 1e-5      # conv_cfg.norm_eps for BatchNorm2d ('rw' variants)

ARCHS = {
    "maxvit_tiny_rw": dict(embed_dim=(64, 128, 256, 512), depths=(2, 2, 5, 2), stem_width=(32, 64), dim_head=32,
                           window=7, expand=4, se_div=16, img=224),
}


def arch_of(model_kind: str) -> dict:
    for k, v in ARCHS.items():
        if k in model_kind.lower():
            return v
    raise ValueError(f"no MaxViT restatement for model_kind {model_kind!r}")


def bn2d(sd, p, x):
    s = sd[p + "weight"] / torch.sqrt(sd[p + "running_var"] + BN_EPS)
    return x * s.view(1, -1, 1, 1) + (sd[p + "bias"] - sd[p + "running_mean"] * s).view(1, -1, 1, 1)


def rel_pos_index(win: int) -> torch.Tensor:
    """Swin-style index into the [(2w-1)^2, heads] bias table for tokens i,j of a row-major w x w window."""
    ys, xs = torch.meshgrid(torch.arange(win), torch.arange(win), indexing="ij")
    y, x = ys.flatten(), xs.flatten()
    return (y[:, None] - y[None, :] + win - 1) * (2 * win - 1) + (x[:, None] - x[None, :] + win - 1)


def window_partition(x, w):          # [B,H,W,C] -> [B*nh*nw, w, w, C]   contiguous w x w tiles
    B, H, W, C = x.shape
    return x.view(B, H // w, w, W // w, w, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, w, w, C)


def window_reverse(t, w, H, W):
    C = t.shape[-1]
    return t.view(-1, H // w, W // w, w, w, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, H, W, C)


def grid_partition(x, g):            # [B,H,W,C] -> [B*(H/g)*(W/g), g, g, C]   g x g tokens strided by H/g
    B, H, W, C = x.shape
    return x.view(B, g, H // g, g, W // g, C).permute(0, 2, 4, 1, 3, 5).reshape(-1, g, g, C)


def grid_reverse(t, g, H, W):
    C = t.shape[-1]
    return t.view(-1, H // g, W // g, g, g, C).permute(0, 3, 1, 4, 2, 5).reshape(-1, H, W, C)


def attention_cl(sd, p, x, dim_head):
    """timm AttentionCl (head_first=True): x [Bw,w,w,C] -> [Bw,w,w,C]."""
    Bw, w, _, C = x.shape
    heads = C // dim_head
    qkv = lin(sd, p + "qkv.", x).view(Bw, w * w, heads, 3 * dim_head).transpose(1, 2)
    q, k, v = qkv.chunk(3, dim=3)
    bias = sd[p + "rel_pos.relative_position_bias_table"][rel_pos_index(w).view(-1)].view(w * w, w * w, heads)
    attn = (q * dim_head ** -0.5) @ k.transpose(-2, -1) + bias.permute(2, 0, 1).unsqueeze(0)
    attn = attn.softmax(dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(Bw, w, w, C)
    return lin(sd, p + "proj.", out)


def partition_attention(sd, p, x, kind, arch):
    """timm PartitionAttentionCl on NHWC: x + attn(LN(x)) (window or grid partitioned); x + MLP(LN(x))."""
    B, H, W, C = x.shape
    w = arch["window"]
    y = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], LN_EPS)
    if kind == "block":
        y = window_reverse(attention_cl(sd, p + "attn.", window_partition(y, w), arch["dim_head"]), w, H, W)
    else:
        y = grid_reverse(attention_cl(sd, p + "attn.", grid_partition(y, w), arch["dim_head"]), w, H, W)
    x = x + y
    y = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], LN_EPS)
    y = lin(sd, p + "mlp.fc2.", F.gelu(lin(sd, p + "mlp.fc1.", y)))
    return x + y


def mbconv(sd, p, x, stride, cin, cout):
    """timm MbConvBlock ('rw': expand from in_chs, stride in the dw conv, SE rd = mid/16, no bias on 1x1 convs)."""
    sc = x
    if stride == 2:
        sc = F.avg_pool2d(sc, 2)
        if cin != cout:
            sc = F.conv2d(sc, sd[p + "shortcut.expand.weight"])
    mid = sd[p + "conv1_1x1.weight"].shape[0]
    y = bn2d(sd, p + "pre_norm.", x)
    y = F.silu(bn2d(sd, p + "norm1.", F.conv2d(y, sd[p + "conv1_1x1.weight"])))
    y = F.conv2d(y, sd[p + "conv2_kxk.weight"], None, stride=stride, padding=1, groups=mid)
    y = F.silu(bn2d(sd, p + "norm2.", y))
    s = y.mean(dim=(2, 3), keepdim=True)
    s = F.silu(F.conv2d(s, sd[p + "se.fc1.weight"], sd[p + "se.fc1.bias"]))
    s = torch.sigmoid(F.conv2d(s, sd[p + "se.fc2.weight"], sd[p + "se.fc2.bias"]))
    y = F.conv2d(y * s, sd[p + "conv3_1x1.weight"])
    return y + sc


def trunk_features(sd: dict, p: str, x: torch.Tensor, arch: dict, capture: dict | None = None) -> torch.Tensor:
    """timm ``forward_features`` incl. the final LayerNorm2d: [B,3,224,224] -> [B,512,7,7]."""
    x = F.conv2d(x, sd[p + "stem.conv1.weight"], None, stride=2, padding=1)
    x = F.silu(bn2d(sd, p + "stem.norm1.", x))
    if capture is not None:
        capture["stem1"] = x
    x = F.conv2d(x, sd[p + "stem.conv2.weight"], None, stride=1, padding=1)
    if capture is not None:
        capture["stem"] = x
    cin = arch["stem_width"][1]
    for i, (c, d) in enumerate(zip(arch["embed_dim"], arch["depths"])):
        for j in range(d):
            q = f"{p}stages.{i}.blocks.{j}."
            x = mbconv(sd, q + "conv.", x, 2 if j == 0 else 1, cin, c)
            cin = c
            if capture is not None:
                capture[f"s{i}b{j}.conv"] = x
            x = x.permute(0, 2, 3, 1)
            x = partition_attention(sd, q + "attn_block.", x, "block", arch)
            if capture is not None:
                capture[f"s{i}b{j}.block"] = x.permute(0, 3, 1, 2)
            x = partition_attention(sd, q + "attn_grid.", x, "grid", arch)
            x = x.permute(0, 3, 1, 2)
            if capture is not None:
                capture[f"s{i}b{j}"] = x
    C = x.shape[1]
    return F.layer_norm(x.permute(0, 2, 3, 1), (C,), sd[p + "norm.weight"], sd[p + "norm.bias"], LN_EPS).permute(0, 3, 1, 2)


def resize(x, arch):
    # architectures.py:44-50 / :90-96
    s = arch["img"]
    if x.shape[-1] != s or x.shape[-2] != s:
        x = F.interpolate(x, size=(s, s), mode="bilinear", align_corners=False)
    return x


def forward(sd: dict, config: dict, image_input=None, metadata_input=None, capture: dict | None = None):
    """Eval-mode logits ``[B,1]`` for ``config['model_name']`` in {MaxViT, mm_MaxViT, frozen_fusion over MaxViT}."""
    name = config["model_name"]
    arch = arch_of(config.get("model_kind", "maxvit_tiny_rw_224.sw_in1k"))
    with torch.no_grad():
        if name == "mm_MaxViT":
            f = trunk_features(sd, "maxvit_backbone.", resize(image_input, arch), arch, capture).mean(dim=(2, 3))
            m = F.gelu(metadata_branch(sd, "metadata_branch.", metadata_input, F.gelu))
            if capture is not None:
                capture["features"], capture["meta"] = f, m
            return head3(sd, "combined_head.", (0, 2, 5), torch.cat((f, m), dim=1), F.gelu)
        if name == "MaxViT":
            f = trunk_features(sd, "maxvit.", resize(image_input, arch), arch, capture).mean(dim=(2, 3))
            if capture is not None:
                capture["features"] = f
            return head3(sd, "maxvit.head.", (1, 3, 6), f, F.gelu)
        if name == "frozen_fusion":
            # architectures.py:296-372 with a MaxViT image branch: head cut to [global_pool] (:304-308), metadata branch
            # um_nn.network[:-2] (pre-activation embedding, :299-303), ReLU fusion head (:357-365)
            icfg = config["image_model_config"]
            arch = arch_of(icfg.get("model_kind", "maxvit_tiny_rw_224.sw_in1k"))
            f = trunk_features(sd, "image_branch.maxvit.", resize(image_input, arch), arch, capture).mean(dim=(2, 3))
            m = metadata_branch(sd, "meta_branch.network.", metadata_input, F.relu)
            if capture is not None:
                capture["features"], capture["meta"] = f, m
            return head3(sd, "combined_head.", (0, 2, 5), torch.cat((f, m), dim=1), F.relu)
    raise ValueError(name)
