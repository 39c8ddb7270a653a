"""Host-side driver of the MaxViT path (timm ``maxvit_tiny_rw_224`` behind `btsbot/architectures.py:25-101`):
weight packing (BatchNorm folding, GEMM layouts) and the kernel sequence of one forward pass.

Like ``_engine`` this file only allocates device memory and calls ``libbtsbot_b200.so``; all arithmetic is in the
CUDA library.  Structure follows SURVEY.md Appendix A.2:

    bilinear 63->224 + stem.conv1(3x3/s2)+BN+SiLU            one fused kernel (the 224x224 image is never materialised)
    stem.conv2 (3x3, 32->64)                                  im2col + tcgen05 GEMM
    per block:  MBConv   = [avgpool2 (+1x1 GEMM)] shortcut; 1x1 GEMM (pre_norm and norm1 BatchNorms folded, SiLU epilogue);
                           dw3x3+BN+SiLU+SE-squeeze; SE gate; gate scaling; 1x1 GEMM with the shortcut added in the epilogue
                attn x2  = LayerNorm; qkv GEMM; window|grid attention (partition = index arithmetic); proj GEMM (+x);
                           LayerNorm; MLP (fused fc1-GELU-fc2 kernel for C <= 160, else two GEMMs) (+x)
    final LayerNorm2d + global average pool                   one kernel -> [B, 512] float32
"""
from __future__ import annotations

import os

import ctypes as C

import torch

from . import _lib as L
from .synth import maxvit_arch

BN2D_EPS = 1e-5     # timm conv_cfg.norm_eps for the 'rw' MaxViT variants

#: images per pass through the trunk (bounds the largest activation: chunk x 112 x 112 x 256 elements)
CHUNK = {"fp32": 128, "bf16": 1024}
#: use the fused fc1->GELU->fc2 kernel in the attention MLPs where it applies (bf16, C <= 160 and C = 256)
FUSE_MLP = True
#: stem.conv2 as an implicit GEMM with 4-D bulk-tensor operand loads; BTSB_CONV3_TC=0 -> im2col3 + GEMM
CONV3_TC = os.environ.get("BTSB_CONV3_TC", "1") != "0"

_DT = {"fp32": (L.F32, torch.float32), "bf16": (L.BF16, torch.bfloat16)}


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


def _bn_fold(sd, p):
    s = sd[p + "weight"].float() / torch.sqrt(sd[p + "running_var"].float() + BN2D_EPS)
    return s, sd[p + "bias"].float() - sd[p + "running_mean"].float() * s


class MaxVitWeights:
    """Kernel-ready copy of a timm-keyed ``maxvit_tiny_rw_224`` trunk."""

    def __init__(self, sd: dict, prefix: str, arch: dict, precision: str):
        code, wdt = _DT[precision]
        self.precision, self.code, self.wdt, self.arch = precision, code, wdt, arch
        g = lambda k: sd[prefix + k].detach()
        dev = g("stem.conv1.weight").device
        sw = arch["stem_width"]
        s1, t1 = _bn_fold(sd, prefix + "stem.norm1.")
        w = g("stem.conv1.weight").float() * s1.view(-1, 1, 1, 1)
        self.stem1_w = w.reshape(sw[0], 27).t().contiguous()                      # [27, C1], k = (ci*3+ky)*3+kx
        self.stem1_shift = t1.contiguous()
        self.stem2_w = g("stem.conv2.weight").permute(0, 2, 3, 1).reshape(sw[1], 9 * sw[0]).to(wdt).contiguous()
        self._zeros = torch.zeros(4 * max(arch["embed_dim"]), device=dev, dtype=torch.float32)
        self._ones = torch.ones(4 * max(arch["embed_dim"]), device=dev, dtype=torch.float32)
        self.blocks = []
        cin = sw[1]
        for i, (c, d) in enumerate(zip(arch["embed_dim"], arch["depths"])):
            for j in range(d):
                q = f"stages.{i}.blocks.{j}."
                blk = dict(cin=cin, c=c, stride=2 if j == 0 else 1, mid=arch["expand"] * cin, name=f"s{i}b{j}")
                m = q + "conv."
                sp, tp = _bn_fold(sd, prefix + m + "pre_norm.")
                sn1, tn1 = _bn_fold(sd, prefix + m + "norm1.")
                w1 = g(m + "conv1_1x1.weight").float().reshape(blk["mid"], cin)
                blk["w1"] = (sn1.view(-1, 1) * w1 * sp.view(1, -1)).to(wdt).contiguous()
                blk["b1"] = (sn1 * (w1 @ tp) + tn1).contiguous()
                sn2, tn2 = _bn_fold(sd, prefix + m + "norm2.")
                blk["dw_w"] = (g(m + "conv2_kxk.weight").float().reshape(blk["mid"], 9) * sn2.view(-1, 1)).t().contiguous()
                blk["dw_shift"] = tn2.contiguous()
                blk["se_w1"] = _f32c(g(m + "se.fc1.weight").reshape(-1, blk["mid"]))
                blk["se_b1"] = _f32c(g(m + "se.fc1.bias"))
                blk["se_w2"] = _f32c(g(m + "se.fc2.weight").reshape(blk["mid"], -1).t())          # [R, mid]
                blk["se_b2"] = _f32c(g(m + "se.fc2.bias"))
                blk["w3"] = g(m + "conv3_1x1.weight").reshape(c, blk["mid"]).to(wdt).contiguous()
                blk["sc_w"] = None
                if blk["stride"] == 2 and cin != c:
                    blk["sc_w"] = g(m + "shortcut.expand.weight").reshape(c, cin).to(wdt).contiguous()
                for part in ("attn_block", "attn_grid"):
                    a = f"{q}{part}."
                    blk[part] = dict(
                        n1_w=_f32c(g(a + "norm1.weight")), n1_b=_f32c(g(a + "norm1.bias")),
                        qkv_w=g(a + "attn.qkv.weight").to(wdt).contiguous(), qkv_b=_f32c(g(a + "attn.qkv.bias")),
                        table=_f32c(g(a + "attn.rel_pos.relative_position_bias_table")),
                        proj_w=g(a + "attn.proj.weight").to(wdt).contiguous(), proj_b=_f32c(g(a + "attn.proj.bias")),
                        n2_w=_f32c(g(a + "norm2.weight")), n2_b=_f32c(g(a + "norm2.bias")),
                        fc1_w=g(a + "mlp.fc1.weight").to(wdt).contiguous(), fc1_b=_f32c(g(a + "mlp.fc1.bias")),
                        fc2_w=g(a + "mlp.fc2.weight").to(wdt).contiguous(), fc2_b=_f32c(g(a + "mlp.fc2.bias")),
                    )
                self.blocks.append(blk)
                cin = c
        self.norm_w, self.norm_b = _f32c(g("norm.weight")), _f32c(g("norm.bias"))
        self.num_features = arch["embed_dim"][-1]

    def zeros(self, n):
        return self._zeros[:n]

    def ones(self, n):
        return self._ones[:n]


def _gemm(name, a, wt, bias, gamma, res, code, epi, st):
    """epi(a @ wt^T + bias) -> new tensor; algorithmic work 2MNK flops, bytes = A + W + out (+ residual)."""
    M, K = a.shape
    N = wt.shape[0]
    out = torch.empty((M, N), device=a.device, dtype=a.dtype)
    es = a.element_size()
    nbytes = es * (M * K + N * K + M * N + (M * N if res is not None else 0)) + 4.0 * N
    L.launch(name, L.lib().btsb_gemm_fwd, _p(a), _p(wt), _p(bias), _p(gamma), _p(res), _p(out), M, N, K, code, epi, st,
             flops=2.0 * M * N * K, nbytes=nbytes)
    return out


def _ln(name, x, w, b, code, st):
    out = torch.empty_like(x)
    M, Cc = x.shape
    L.launch(name, L.lib().btsb_layernorm_rows_fwd, _p(x), _p(w), _p(b), _p(out), M, Cc, code, st,
             flops=8.0 * x.numel(), nbytes=2.0 * x.element_size() * x.numel())
    return out


def _attention_block(w: MaxVitWeights, a: dict, cur, B, H, W, c, grid_mode, tag, st, capture, capname):
    lib, code = L.lib(), w.code
    es = cur.element_size()
    M = B * H * W
    y = _ln(f"mv_ln_{c}", cur, a["n1_w"], a["n1_b"], code, st)
    qkv = _gemm(f"mv_qkv_{c}", y, a["qkv_w"], a["qkv_b"], None, None, code, L.EPI_BIAS, st)
    o = torch.empty((M, c), device=cur.device, dtype=cur.dtype)
    heads = c // 32
    L.launch(f"mv_attn_{tag}_{c}", lib.btsb_maxvit_attn_fwd, _p(qkv), _p(o), B, H, W, c, grid_mode, _p(a["table"]), code, st,
             flops=4.0 * 49 * 32 * M * heads, nbytes=es * (qkv.numel() + o.numel()))
    cur = _gemm(f"mv_proj_{c}", o, a["proj_w"], a["proj_b"], w.ones(c), cur, code, L.EPI_SCALE_RES, st)
    y = _ln(f"mv_ln_{c}", cur, a["n2_w"], a["n2_b"], code, st)
    if FUSE_MLP and code == L.BF16 and c % 16 == 0 and (64 <= c <= 160 or c == 256):
        nxt = torch.empty_like(cur)
        L.launch(f"mv_mlp_fused_{c}", lib.btsb_convnext_mlp_fused_fwd, _p(y), _p(cur), _p(a["fc1_w"]), _p(a["fc1_b"]),
                 _p(a["fc2_w"]), _p(a["fc2_b"]), _p(w.ones(c)), _p(nxt), M, c, L.BF16, st,
                 flops=16.0 * M * c * c, nbytes=es * (3.0 * M * c + 8.0 * c * c))
        cur = nxt
    else:
        hid = _gemm(f"mv_fc1_{c}", y, a["fc1_w"], a["fc1_b"], None, None, code, L.EPI_BIAS_GELU, st)
        cur = _gemm(f"mv_fc2_{c}", hid, a["fc2_w"], a["fc2_b"], w.ones(c), cur, code, L.EPI_SCALE_RES, st)
    if capture is not None:
        capture[capname] = (cur, H, W)
    return cur


def _trunk_chunk(w: MaxVitWeights, x: torch.Tensor, capture: dict | None, feat: torch.Tensor):
    lib, code, adt = L.lib(), w.code, w.wdt
    arch = w.arch
    B, _, Hin, Win = x.shape
    S = arch["img"]
    dev, st = x.device, L.stream_ptr()
    es = 2 if adt == torch.bfloat16 else 4
    sw = arch["stem_width"]
    H = W = S // 2
    a = torch.empty((B * H * W, sw[0]), device=dev, dtype=adt)
    L.launch("mv_stem1", lib.btsb_maxvit_stem1_fwd, _p(x), B, Hin, Win, S, _p(w.stem1_w), _p(w.stem1_shift), sw[0],
             _p(a), code, st, flops=2.0 * 27 * sw[0] * B * H * W, nbytes=4.0 * x.numel() + es * a.numel())
    if capture is not None:
        capture["stem1"] = (a, H, W)
    if CONV3_TC and code == L.BF16 and sw[0] == 32 and sw[1] == 64 and H % 8 == 0 and W % 16 == 0:
        # implicit GEMM: every tap's A operand is one 4-D bulk tensor copy of the shifted input box (no patch matrix)
        cur = torch.empty((B * H * W, sw[1]), device=dev, dtype=adt)
        L.launch("mv_stem2_conv", lib.btsb_conv3x3_c32_fwd, _p(a), _p(w.stem2_w), None, _p(cur), B, H, W, sw[1], st,
                 flops=2.0 * 9 * sw[0] * sw[1] * B * H * W, nbytes=es * (a.numel() + cur.numel()))
        del a
    else:
        col = torch.empty((B * H * W, 9 * sw[0]), device=dev, dtype=adt)
        L.launch("mv_im2col3", lib.btsb_maxvit_im2col3_fwd, _p(a), _p(col), B, H, W, sw[0], code, st,
                 nbytes=es * (a.numel() + col.numel()))
        cur = _gemm("mv_stem2", col, w.stem2_w, w.zeros(sw[1]), None, None, code, L.EPI_BIAS, st)
        del col, a
    if capture is not None:
        capture["stem"] = (cur, H, W)
    for blk in w.blocks:
        cin, c, mid, stride = blk["cin"], blk["c"], blk["mid"], blk["stride"]
        Ho, Wo = H // stride, W // stride
        # ---- MBConv -------------------------------------------------------------------------------------
        if stride == 2:
            sc = torch.empty((B * Ho * Wo, cin), device=dev, dtype=adt)
            L.launch("mv_avgpool2", lib.btsb_maxvit_avgpool2_fwd, _p(cur), _p(sc), B, H, W, cin, code, st,
                     nbytes=es * (cur.numel() + sc.numel()))
            if blk["sc_w"] is not None:
                sc = _gemm(f"mv_shortcut_{c}", sc, blk["sc_w"], w.zeros(c), None, None, code, L.EPI_BIAS, st)
        else:
            sc = cur
        h1 = _gemm(f"mv_expand_{cin}", cur, blk["w1"], blk["b1"], None, None, code, L.EPI_BIAS_SILU, st)
        h2 = torch.empty((B * Ho * Wo, mid), device=dev, dtype=adt)
        pooled = torch.empty((B, mid), device=dev, dtype=torch.float32)
        L.launch(f"mv_dw3_s{stride}_{mid}", lib.btsb_maxvit_dw3_fwd, _p(h1), B, H, W, mid, stride, _p(blk["dw_w"]),
                 _p(blk["dw_shift"]), _p(h2), _p(pooled), code, st,
                 flops=2.0 * 9 * h2.numel(), nbytes=es * (h1.numel() + h2.numel()))
        del h1
        gate = torch.empty((B, mid), device=dev, dtype=torch.float32)
        rd = blk["se_w1"].shape[0]
        L.launch("mv_se", lib.btsb_maxvit_se_fwd, _p(pooled), B, mid, rd, _p(blk["se_w1"]), _p(blk["se_b1"]),
                 _p(blk["se_w2"]), _p(blk["se_b2"]), _p(gate), st, flops=4.0 * B * mid * rd, nbytes=8.0 * B * mid)
        L.launch(f"mv_scale_{mid}", lib.btsb_maxvit_scale_fwd, _p(h2), _p(gate), B, Ho * Wo, mid, code, st,
                 flops=1.0 * h2.numel(), nbytes=2.0 * es * h2.numel())
        cur = _gemm(f"mv_project_{c}", h2, blk["w3"], w.zeros(c), w.ones(c), sc, code, L.EPI_SCALE_RES, st)
        del h2, sc
        H, W = Ho, Wo
        if capture is not None:
            capture[blk["name"] + ".conv"] = (cur, H, W)
        # ---- window attention, then grid attention ------------------------------------------------------------
        cur = _attention_block(w, blk["attn_block"], cur, B, H, W, c, 0, "win", st, capture, blk["name"] + ".block")
        cur = _attention_block(w, blk["attn_grid"], cur, B, H, W, c, 1, "grid", st, capture, blk["name"])
    L.launch("mv_lnpool", lib.btsb_maxvit_lnpool_fwd, _p(cur), _p(w.norm_w), _p(w.norm_b), _p(feat), B, H * W,
             w.num_features, code, st, flops=8.0 * cur.numel(), nbytes=es * cur.numel() + 4.0 * feat.numel())


def maxvit_features(w: MaxVitWeights, x: torch.Tensor, capture: dict | None = None, chunk: int | None = None) -> torch.Tensor:
    """``F.interpolate`` to 224 (when needed) + timm ``forward_features`` + final norm + global pool -> ``[B, 512]`` float32.
    ``x``: [B,3,H,W] float32 CUDA (architectures.py:42-51,88-97)."""
    L.require_cuda(x, "image input")
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError(f"expected image batch [B,3,H,W], got {tuple(x.shape)}")
    x = x.to(torch.float32).contiguous()
    B = x.shape[0]
    chunk = chunk or CHUNK[w.precision]
    if capture is not None and B > chunk:
        raise ValueError("capture needs the batch to fit one chunk")
    out = torch.empty((B, w.num_features), device=x.device, dtype=torch.float32)
    for lo in range(0, B, chunk):         # row slices of contiguous tensors: views, no copies
        _trunk_chunk(w, x[lo:lo + chunk], capture, out[lo:lo + chunk])
    return out


def arch_for(config: dict) -> dict:
    return maxvit_arch(config.get("model_kind", "maxvit_tiny_rw_224.sw_in1k"))
