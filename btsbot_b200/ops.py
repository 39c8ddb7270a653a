"""Per-kernel Python entry points (torch CUDA tensors in/out) over the C ABI -- used by the unit tests, the
training path and anyone who wants a single fused op.  Shapes/dtypes: see ``include/btsbot_b200.h``."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

#: tensor dtype of the call's residual-stream operand -> dtype code (float16 rows: bf16 compute, fp16 residual stream)
_CODE = {torch.float32: L.F32, torch.bfloat16: L.BF16, torch.float64: L.F64, torch.float16: L.BF16_XF16}
_OUT = {torch.float32: torch.float32, torch.bfloat16: torch.bfloat16, torch.float16: torch.bfloat16}


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _chk(*ts):
    for t in ts:
        if t is not None:
            L.require_cuda(t, "tensor")
            if not t.is_contiguous():
                raise ValueError("btsbot_b200.ops: tensors must be contiguous")


def stem(x, w48, bias, ln_w, ln_b, out_dtype=torch.float32):
    """x [B,3,H,W] f32; w48 [48,C0] f32 -> rows [B*h*w, C0]."""
    _chk(x, w48, bias, ln_w, ln_b)
    B, _, H, W = x.shape
    c0 = w48.shape[1]
    h, w = (H - 4) // 4 + 1, (W - 4) // 4 + 1
    out = torch.empty((B * h * w, c0), device=x.device, dtype=out_dtype)
    L.check(L.lib().btsb_convnext_stem_fwd(_p(x), B, H, W, _p(w48), _p(bias), _p(ln_w), _p(ln_b), c0, _p(out),
                                           _CODE[out_dtype], L.stream_ptr()), "stem")
    return out


def dwln(x, B, H, W, w49, bias, ln_w, ln_b):
    """x rows [B*H*W, C] (f32|bf16); w49 [49,C] f32 -> same shape/dtype.  fp16 rows (the residual stream of the bf16
    mode) give bf16 output rows."""
    _chk(x, w49, bias, ln_w, ln_b)
    out = torch.empty_like(x, dtype=_OUT[x.dtype])
    L.check(L.lib().btsb_convnext_dwln_fwd(_p(x), _CODE[x.dtype], B, H, W, x.shape[1], _p(w49), _p(bias), _p(ln_w),
                                           _p(ln_b), _p(out), L.stream_ptr()), "dwln")
    return out


def lnpatch(x, B, H, W, ln_w, ln_b):
    _chk(x, ln_w, ln_b)
    c = x.shape[1]
    ho, wo = (H - 2) // 2 + 1, (W - 2) // 2 + 1
    out = torch.empty((B * ho * wo, 4 * c), device=x.device, dtype=_OUT[x.dtype])     # fp16 stream rows in -> bf16 out
    L.check(L.lib().btsb_convnext_lnpatch_fwd(_p(x), _CODE[x.dtype], B, H, W, c, _p(ln_w), _p(ln_b), _p(out),
                                              L.stream_ptr()), "lnpatch")
    return out


def poolln(x, B, HW, ln_w=None, ln_b=None):
    _chk(x, ln_w, ln_b)
    out = torch.empty((B, x.shape[1]), device=x.device, dtype=torch.float32)
    L.check(L.lib().btsb_convnext_poolln_fwd(_p(x), _CODE[x.dtype], B, HW, x.shape[1], _p(ln_w), _p(ln_b), _p(out),
                                             L.stream_ptr()), "poolln")
    return out


def gemm(a, wt, bias, epilogue=L.EPI_BIAS, gamma=None, res=None, out_dtype=None):
    """out[M,N] = epi(a[M,K] @ wt[N,K]^T + bias); a/wt/res share dtype (f32 -> CUDA cores, bf16 -> tcgen05).
    bf16 operands with ``out_dtype=torch.float16`` (and an fp16 ``res``): the output joins the fp16 residual stream."""
    _chk(a, wt, bias, gamma, res)
    xf16 = a.dtype == torch.bfloat16 and (out_dtype == torch.float16 or (res is not None and res.dtype == torch.float16))
    if a.dtype != wt.dtype or (res is not None and res.dtype != (torch.float16 if xf16 else a.dtype)):
        raise ValueError("gemm: a, wt and res must share one dtype (res fp16 with bf16 operands: fp16 output)")
    M, K = a.shape
    N = wt.shape[0]
    out = torch.empty((M, N), device=a.device, dtype=torch.float16 if xf16 else a.dtype)
    L.check(L.lib().btsb_gemm_fwd(_p(a), _p(wt), _p(bias), _p(gamma), _p(res), _p(out), M, N, K,
                                  L.BF16_XF16 if xf16 else _CODE[a.dtype], epilogue, L.stream_ptr()), "gemm")
    return out


def mlp_fused(y, res, w1, b1, w2, b2, gamma, inplace=False):
    """res + gamma * (fc2(gelu(fc1(y)+b1))+b2) in one tcgen05 kernel (bf16, C in [64,160], 256, 320); ``res`` bf16 or
    fp16 (the residual stream), the result has its dtype.
    ``inplace=True`` passes out == res (the wide variants then add the update to ``res`` with a bulk tensor reduction)."""
    _chk(y, res, w1, b1, w2, b2, gamma)
    M, c = y.shape
    out = res if inplace else torch.empty_like(res)
    L.check(L.lib().btsb_convnext_mlp_fused_fwd(_p(y), _p(res), _p(w1), _p(b1), _p(w2), _p(b2), _p(gamma), _p(out),
                                                M, c, _CODE[res.dtype], L.stream_ptr()), "mlp_fused")
    return out


def stem_tc(x, w_pad, bias, ln_w, ln_b):
    """bf16 tensor-core stem: im2col (K padded 48->64) + tcgen05 GEMM with bias+LayerNorm epilogue.
    x [B,3,H,W] f32; w_pad [C0,64] bf16 -> rows [B*h*w, C0] bf16."""
    _chk(x, w_pad, bias, ln_w, ln_b)
    B, _, H, W = x.shape
    h, w = (H - 4) // 4 + 1, (W - 4) // 4 + 1
    c0 = w_pad.shape[0]
    patches = torch.empty((B * h * w, 64), device=x.device, dtype=torch.bfloat16)
    L.check(L.lib().btsb_stem_im2col_bf16(_p(x), _p(patches), B, H, W, L.stream_ptr()), "im2col")
    out = torch.empty((B * h * w, c0), device=x.device, dtype=torch.bfloat16)
    L.check(L.lib().btsb_gemm_ln_fwd(_p(patches), _p(w_pad), _p(bias), _p(ln_w), _p(ln_b), _p(out), B * h * w, c0, 64,
                                     L.stream_ptr()), "gemm_ln")
    return out, patches


def stem_fused(x, w_pad, bias, ln_w, ln_b, out_dtype=torch.bfloat16):
    """one-kernel bf16 stem (producer warps build the im2col rows in shared memory): rows [B*h*w, C0] bf16, or fp16
    (``out_dtype=torch.float16``: the rows open the fp16 residual stream)."""
    _chk(x, w_pad, bias, ln_w, ln_b)
    B, _, H, W = x.shape
    h, w = (H - 4) // 4 + 1, (W - 4) // 4 + 1
    c0 = w_pad.shape[0]
    out = torch.empty((B * h * w, c0), device=x.device, dtype=out_dtype)
    L.check(L.lib().btsb_stem_fused_fwd(_p(x), B, H, W, _p(w_pad), _p(bias), _p(ln_w), _p(ln_b), _p(out), c0,
                                        _CODE[out_dtype], L.stream_ptr()), "stem_fused")
    return out


def score(logits):
    """sigmoid + 0.5 threshold on device -> (scores f32, labels uint8)."""
    _chk(logits)
    flat = logits.reshape(-1)
    s = torch.empty_like(flat)
    lab = torch.empty(flat.shape, device=flat.device, dtype=torch.uint8)
    L.check(L.lib().btsb_score_epilogue(_p(flat), flat.numel(), _p(s), _p(lab), L.stream_ptr()), "score")
    return s.view_as(logits), lab.view_as(logits)
