"""Training-mode forward + hand-written backward of the reference models on the CUDA kernels (K7): all-fp32
(``precision="fp32"``, the reference-numerics path) or mixed precision (``precision="bf16"``: the fc1 / fc2 /
downsample GEMMs -- forward, dgrad and split-K wgrad (MN-major operands: no transposed activation copies) -- on tcgen05
tensor cores with bf16 operands and fp32 accumulation, everything else in fp32; the arithmetic torch autocast(bf16)
would give train.py).

Replaces, for one step of `btsbot/train.py:496-547`, everything PyTorch autograd + cuDNN/ATen would do between
``logits = model(...)`` and ``loss.backward()``: the forward saves the intermediates each backward kernel needs and
the backward writes parameter gradients (dgrad/wgrad GEMMs, depthwise-conv and LayerNorm backward, BatchNorm1d with
batch statistics, dropout masks).  ``torch.autograd`` only sees one opaque ``Function`` whose input is the logits'
gradient.  Gradients land in ``param.grad`` (accumulating, like autograd) and -- when a gradient sink is installed by
``btsbot_b200.parallel`` -- are handed over bucket by bucket so the NCCL all-reduce overlaps the rest of the backward.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from .synth import convnext_arch

LN_EPS = 1e-6
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _st():
    return L.stream_ptr()


def _new(shape, like):
    return torch.empty(shape, device=like.device, dtype=torch.float32)


# ---- thin op layer (every call is one kernel of libbtsbot_b200.so) ------------------------------------------
def gemm_nt(a, w, bias=None):
    """a [M,K] @ w[N,K]^T (+ bias)."""
    M, K = a.shape
    N = w.shape[0]
    out = _new((M, N), a)
    L.launch("t_gemm_nt", L.lib().btsb_gemm_f32_strided, _p(a), K, 1, _p(w), 1, K, _p(out), M, N, K, 0, _st(),
             flops=2.0 * M * N * K)
    if bias is not None:
        L.launch("t_bias", L.lib().btsb_bias_add_f32, _p(out), _p(bias), M, N, _st())
    return out


def gemm_nn(a, w):
    """a [M,N] @ w[N,K] -> [M,K]   (dgrad)."""
    M, N = a.shape
    K = w.shape[1]
    out = _new((M, K), a)
    L.launch("t_gemm_nn", L.lib().btsb_gemm_f32_strided, _p(a), N, 1, _p(w), K, 1, _p(out), M, K, N, 0, _st(),
             flops=2.0 * M * N * K)
    return out


def gemm_tn(a, b):
    """a[M,N]^T @ b[M,K] -> [N,K]   (wgrad; split-K inside)."""
    M, N = a.shape
    K = b.shape[1]
    out = _new((N, K), a)
    L.launch("t_gemm_tn", L.lib().btsb_gemm_f32_strided, _p(a), 1, N, _p(b), K, 1, _p(out), N, K, M, 0, _st(),
             flops=2.0 * M * N * K)
    return out


# ---- bf16 tensor-core GEMMs (mixed-precision mode: fp32 residual stream / LayerNorm / element-wise, bf16 operands) ---
def _ld(M):
    return (M + 7) // 8 * 8


def cast_dual(x, rm=True, t=True, op=0, x2=None, colvec=None, want_colsum=False, aux=None):
    """One pass over fp32 ``x`` [M,N]: bf16 row-major copy, bf16 transposed copy [N, ld], optional fused GELU (op 1) /
    GELU backward (op 2, x = pre-activation, x2 = upstream gradient), column scale and column sums."""
    M, N = x.shape
    ld = _ld(M)
    in16 = x.dtype == torch.bfloat16
    assert x2 is None or x2.dtype == x.dtype
    o_rm = torch.empty((M, N), device=x.device, dtype=torch.bfloat16) if rm else None
    o_t = torch.empty((N, ld), device=x.device, dtype=torch.bfloat16) if t else None
    cs = torch.zeros((N,), device=x.device, dtype=torch.float32) if want_colsum else None
    eb = 2 if in16 else 4
    asum = torch.zeros((N,), device=x.device, dtype=torch.float32) if aux is not None else None
    L.launch("t_cast_dual", L.lib().btsb_cast_dual_bf16, _p(x), _p(x2), _p(colvec), _p(o_rm), _p(o_t), _p(cs), M, N, ld,
             op, L.BF16 if in16 else L.F32, _p(aux), _p(asum), _st(),
             nbytes=float(x.numel()) * (eb + (eb if x2 is not None else 0) + (2 if rm else 0) + (2 if t else 0) +
                                        (4 if aux is not None else 0)))
    if aux is not None:
        return o_rm, o_t, cs, asum
    return o_rm, o_t, cs


def tc_gemm(a16, w16, bias=None):
    """fp32 [M,N] = a16 [M,K] @ w16 [N,K]^T (+ bias) on tcgen05."""
    M, K = a16.shape
    N = w16.shape[0]
    out = torch.empty((M, N), device=a16.device, dtype=torch.float32)
    L.launch("t_gemm_tc", L.lib().btsb_gemm_bf16_f32out, _p(a16), _p(w16), _p(bias), _p(out), M, N, K, _st(),
             flops=2.0 * M * N * K, nbytes=2.0 * (M * K + N * K) + 4.0 * M * N)
    return out


def tc_gemm16(a16, w16, bias):
    """bf16 [M,N] = a16 [M,K] @ w16 [N,K]^T + bias through the inference GEMM (bulk-tensor-store epilogue): used for the
    4C-wide tensors (fc1 pre-activation, fc2 dgrad), which autocast would keep in bf16 as well."""
    M, K = a16.shape
    N = w16.shape[0]
    out = torch.empty((M, N), device=a16.device, dtype=torch.bfloat16)
    L.launch("t_gemm_tc16", L.lib().btsb_gemm_fwd, _p(a16), _p(w16), _p(bias), None, None, _p(out), M, N, K, L.BF16,
             L.EPI_BIAS, _st(), flops=2.0 * M * N * K, nbytes=2.0 * (M * K + N * K + M * N))
    return out


def tc_wgrad(at16, bt16, K):
    """fp32 [Mo,No] = at16 [Mo, ld] (first K columns) @ bt16 [No, ld]^T: contraction over the activation rows."""
    Mo, ld = at16.shape
    No = bt16.shape[0]
    out = torch.zeros((Mo, No), device=at16.device, dtype=torch.float32)
    L.launch("t_wgrad_tc", L.lib().btsb_gemm_bf16_wgrad, _p(at16), _p(bt16), ld, _p(out), Mo, No, K, _st(),
             flops=2.0 * Mo * No * K, nbytes=2.0 * K * (Mo + No) + 4.0 * Mo * No)
    return out


def tc_wgrad_mn(a16, b16):
    """fp32 [Mo,No] = a16 [K, Mo]^T @ b16 [K, No] with both operands ROW-MAJOR bf16 (MN-major UMMA operands): the wgrad
    GEMM reads dY and X as the forward / dgrad GEMMs do, no transposed copies."""
    K, Mo = a16.shape
    No = b16.shape[1]
    out = torch.zeros((Mo, No), device=a16.device, dtype=torch.float32)
    L.launch("t_wgrad_tc", L.lib().btsb_gemm_bf16_wgrad_mn, _p(a16), _p(b16), _p(out), Mo, No, K, _st(),
             flops=2.0 * Mo * No * K, nbytes=2.0 * K * (Mo + No) + 4.0 * Mo * No)
    return out


def _w16(w2d):
    """bf16 copies of a weight matrix [N,K]: (row-major [N,K] for the forward, transposed [K,N] for dgrad)."""
    N, K = w2d.shape
    assert N % 8 == 0
    rm, t, _ = cast_dual(w2d.contiguous())
    return rm, t            # t is [K, ld = N]


def colsum(x, y=None):
    M, N = x.shape
    out = _new((N,), x)
    L.launch("t_colsum", L.lib().btsb_colsum_f32, _p(x), _p(y), _p(out), M, N, 0, _st())
    return out


def act(pre, kind, dout=None):
    out = torch.empty_like(pre)
    L.launch("t_act", L.lib().btsb_act_f32, _p(pre), _p(dout), _p(out), pre.numel(), kind, _st())
    return out


def colscale(x, g, res=None):
    M, N = x.shape
    out = torch.empty_like(x)
    L.launch("t_colscale", L.lib().btsb_colscale_f32, _p(x), _p(g), _p(res), _p(out), M, N, _st())
    return out


def ln_fwd(u, w, b):
    y = torch.empty_like(u)
    L.launch("t_ln_fwd", L.lib().btsb_layernorm_fwd_f32, _p(u), _p(w), _p(b), _p(y), u.shape[0], u.shape[1], LN_EPS, _st())
    return y


def ln_bwd(u, w, dy, need_du=True):
    du = torch.empty_like(u) if need_du else None
    dw = torch.zeros_like(w)
    db = torch.zeros_like(w)
    L.launch("t_ln_bwd", L.lib().btsb_layernorm_bwd_f32, _p(u), _p(w), _p(dy), _p(du), _p(dw), _p(db), u.shape[0],
             u.shape[1], LN_EPS, _st())
    return du, dw, db


def dwconv(x, w49, bias, B, H, W, flip=0):
    out = torch.empty_like(x)
    L.launch("t_dwconv", L.lib().btsb_dwconv7_f32, _p(x), _p(w49), _p(bias), _p(out), B, H, W, x.shape[1], flip, _st(),
             flops=98.0 * x.numel())
    return out


def dwconv_wgrad(x, du, B, H, W):
    c = x.shape[1]
    dw = torch.zeros((49, c), device=x.device, dtype=torch.float32)
    db = torch.zeros((c,), device=x.device, dtype=torch.float32)
    L.launch("t_dwconv_wgrad", L.lib().btsb_dwconv7_wgrad_f32, _p(x), _p(du), _p(dw), _p(db), B, H, W, c, _st(),
             flops=98.0 * x.numel())
    return dw, db


def patch2x2(src, B, H, W, c, reverse):
    ho, wo = (H - 2) // 2 + 1, (W - 2) // 2 + 1
    dst = _new((B * H * W, c), src) if reverse else _new((B * ho * wo, 4 * c), src)
    L.launch("t_patch", L.lib().btsb_patch2x2_f32, _p(src), _p(dst), B, H, W, c, int(reverse), _st())
    return dst


def pool(src, B, HW, c, reverse):
    dst = _new((B * HW, c), src) if reverse else _new((B, c), src)
    L.launch("t_pool", L.lib().btsb_pool_f32, _p(src), _p(dst), B, HW, c, int(reverse), _st())
    return dst


def dropout(x, p, mask=None, seed=0, counter=None):
    """``counter``: optional device int64 tensor mixed into the seed inside the kernel (CUDA-graph replays keep the
    host-side ``seed`` of the captured launch; the counter, bumped once per forward, still changes the mask)."""
    y = torch.empty_like(x)
    reuse = mask is not None
    if mask is None:
        mask = torch.empty(x.shape, device=x.device, dtype=torch.uint8)
    if counter is not None and not reuse:
        L.launch("t_dropout", L.lib().btsb_dropout_ctr_f32, _p(x), _p(y), _p(mask), x.numel(), float(p), int(seed),
                 _p(counter), 0, _st())
    else:
        L.launch("t_dropout", L.lib().btsb_dropout_f32, _p(x), _p(y), _p(mask), x.numel(), float(p), int(seed), int(reuse), _st())
    return y, mask


def _seed():
    return int(torch.randint(0, 2 ** 62, (1,)).item())         # CPU generator: follows torch.manual_seed


# ---- gradient bookkeeping ----------------------------------------------------------------------------------------
class _Grads:
    """Accumulates into ``param.grad`` and reports finished parameters to an optional sink (DDP buckets)."""

    def __init__(self, sink=None):
        self.sink = sink

    def put(self, param, g):
        if param is None or not param.requires_grad:
            return
        g = g.reshape(param.shape)
        if param.grad is None:
            param.grad = g.contiguous() if self.sink is None else self.sink.adopt(param, g)
        else:
            param.grad.add_(g)
        if self.sink is not None:
            self.sink.ready(param)


def _needs(*params):
    return any(p is not None and p.requires_grad for p in params)


# ---- ConvNeXt trunk ----------------------------------------------------------------------------------------------
def _trunk_fwd(tr, x):
    """tr: ConvNeXtTrunk module.  Returns (rows, h, w, tape)."""
    B, _, H, W = x.shape
    arch = tr.arch
    dims, depths = arch["dims"], arch["depths"]
    h, w = (H - 4) // 4 + 1, (W - 4) // 4 + 1
    c0 = dims[0]
    patches = _new((B * h * w, 48), x)
    L.launch("t_im2col", L.lib().btsb_stem_im2col_f32, _p(x), _p(patches), B, H, W, _st())
    stem_c, stem_n = tr.stem[0], tr.stem[1]
    u0 = gemm_nt(patches, stem_c.weight.detach().reshape(c0, 48), stem_c.bias.detach())
    cur = ln_fwd(u0, stem_n.weight.detach(), stem_n.bias.detach())
    tape = {"B": B, "stem": (patches, u0), "stages": []}
    for i, stage in enumerate(tr.stages):
        c = dims[i]
        st = {"blocks": []}
        if i > 0:
            cin = dims[i - 1]
            ln_m, conv_m = stage.downsample[0], stage.downsample[1]
            yln = ln_fwd(cur, ln_m.weight.detach(), ln_m.bias.detach())
            pt = patch2x2(yln, B, h, w, cin, reverse=False)
            wds = conv_m.weight.detach().permute(0, 2, 3, 1).reshape(c, 4 * cin).contiguous()
            st["down"] = (cur, pt, wds, h, w)
            h, w = (h - 2) // 2 + 1, (w - 2) // 2 + 1
            cur = gemm_nt(pt, wds, conv_m.bias.detach())
        for blk in stage.blocks:
            w49 = blk.conv_dw.weight.detach().reshape(c, 49).t().contiguous()
            u = dwconv(cur, w49, blk.conv_dw.bias.detach(), B, h, w)
            y = ln_fwd(u, blk.norm.weight.detach(), blk.norm.bias.detach())
            w1 = blk.mlp.fc1.weight.detach().reshape(4 * c, c)
            w2 = blk.mlp.fc2.weight.detach().reshape(c, 4 * c)
            hp = gemm_nt(y, w1, blk.mlp.fc1.bias.detach())
            hh = act(hp, L.ACT_GELU)
            v = gemm_nt(hh, w2, blk.mlp.fc2.bias.detach())
            out = colscale(v, blk.gamma.detach(), res=cur)
            st["blocks"].append((cur, u, y, hp, hh, v, w49, h, w))
            cur = out
        tape["stages"].append(st)
    return cur, h, w, tape


def _trunk_bwd(tr, tape, dcur, G):
    B = tape["B"]
    dims = tr.arch["dims"]
    for i in reversed(range(len(tr.stages))):
        stage, st, c = tr.stages[i], tape["stages"][i], dims[i]
        ones = torch.ones((c,), device=dcur.device, dtype=torch.float32)
        for blk, saved in zip(reversed(list(stage.blocks)), reversed(st["blocks"])):
            xin, u, y, hp, hh, v, w49, h, w = saved
            G.put(blk.gamma, colsum(dcur, v))
            dv = colscale(dcur, blk.gamma.detach())
            G.put(blk.mlp.fc2.weight, gemm_tn(dv, hh))
            G.put(blk.mlp.fc2.bias, colsum(dv))
            dhh = gemm_nn(dv, blk.mlp.fc2.weight.detach().reshape(c, 4 * c))
            dhp = act(hp, L.ACT_GELU, dout=dhh)
            G.put(blk.mlp.fc1.weight, gemm_tn(dhp, y))
            G.put(blk.mlp.fc1.bias, colsum(dhp))
            dy = gemm_nn(dhp, blk.mlp.fc1.weight.detach().reshape(4 * c, c))
            du, dlw, dlb = ln_bwd(u, blk.norm.weight.detach(), dy)
            G.put(blk.norm.weight, dlw)
            G.put(blk.norm.bias, dlb)
            dw49, dbias = dwconv_wgrad(xin, du, B, h, w)
            G.put(blk.conv_dw.weight, dw49.t().contiguous())
            G.put(blk.conv_dw.bias, dbias)
            dconv = dwconv(du, w49, None, B, h, w, flip=1)
            dcur = colscale(dconv, ones, res=dcur)                 # conv path + shortcut
        if i > 0:
            xin, pt, wds, h, w = st["down"]
            cin = dims[i - 1]
            ln_m, conv_m = stage.downsample[0], stage.downsample[1]
            dwds = gemm_tn(dcur, pt)                               # [c, (dy,dx,cin)]
            G.put(conv_m.weight, dwds.view(c, 2, 2, cin).permute(0, 3, 1, 2).contiguous())
            G.put(conv_m.bias, colsum(dcur))
            dpt = gemm_nn(dcur, wds)
            dyln = patch2x2(dpt, B, h, w, cin, reverse=True)
            dcur, dlw, dlb = ln_bwd(xin, ln_m.weight.detach(), dyln)
            G.put(ln_m.weight, dlw)
            G.put(ln_m.bias, dlb)
    patches, u0 = tape["stem"]
    stem_c, stem_n = tr.stem[0], tr.stem[1]
    du0, dlw, dlb = ln_bwd(u0, stem_n.weight.detach(), dcur)
    G.put(stem_n.weight, dlw)
    G.put(stem_n.bias, dlb)
    G.put(stem_c.weight, gemm_tn(du0, patches))
    G.put(stem_c.bias, colsum(du0))


def _trunk_fwd_tc(tr, x):
    """Mixed-precision twin of :func:`_trunk_fwd`: the fc1 / fc2 / downsample GEMMs run on the tensor cores (bf16
    operands, fp32 accumulate and fp32 results); the hidden activation gelu(fc1) only ever exists in bf16.  The stem
    GEMM (K = 48, 0.7 % of the FLOPs) stays on the fp32 kernel."""
    B, _, H, W = x.shape
    arch = tr.arch
    dims = arch["dims"]
    h, w = (H - 4) // 4 + 1, (W - 4) // 4 + 1
    c0 = dims[0]
    patches = _new((B * h * w, 48), x)
    L.launch("t_im2col", L.lib().btsb_stem_im2col_f32, _p(x), _p(patches), B, H, W, _st())
    stem_c, stem_n = tr.stem[0], tr.stem[1]
    u0 = gemm_nt(patches, stem_c.weight.detach().reshape(c0, 48), stem_c.bias.detach())
    cur = ln_fwd(u0, stem_n.weight.detach(), stem_n.bias.detach())
    tape = {"B": B, "stem": (patches, u0), "stages": [], "tc": True}
    for i, stage in enumerate(tr.stages):
        c = dims[i]
        st = {"blocks": []}
        if i > 0:
            cin = dims[i - 1]
            ln_m, conv_m = stage.downsample[0], stage.downsample[1]
            yln = ln_fwd(cur, ln_m.weight.detach(), ln_m.bias.detach())
            pt = patch2x2(yln, B, h, w, cin, reverse=False)
            wds = conv_m.weight.detach().permute(0, 2, 3, 1).reshape(c, 4 * cin).contiguous()
            wds16, wds16t = _w16(wds)
            pt16, _, _ = cast_dual(pt, t=False)
            st["down"] = (cur, pt16, wds16t, h, w)
            h, w = (h - 2) // 2 + 1, (w - 2) // 2 + 1
            cur = tc_gemm(pt16, wds16, conv_m.bias.detach())
        for blk in stage.blocks:
            w49 = blk.conv_dw.weight.detach().reshape(c, 49).t().contiguous()
            u = dwconv(cur, w49, blk.conv_dw.bias.detach(), B, h, w)
            y = ln_fwd(u, blk.norm.weight.detach(), blk.norm.bias.detach())
            w1_16, w1_16t = _w16(blk.mlp.fc1.weight.detach().reshape(4 * c, c))
            w2_16, w2_16t = _w16(blk.mlp.fc2.weight.detach().reshape(c, 4 * c))
            y16, _, _ = cast_dual(y, t=False)
            hp = tc_gemm16(y16, w1_16, blk.mlp.fc1.bias.detach())         # bf16 [M, 4c]
            hh16, _, _ = cast_dual(hp, t=False, op=1)                    # gelu fused into the cast
            v = tc_gemm(hh16, w2_16, blk.mlp.fc2.bias.detach())
            out = colscale(v, blk.gamma.detach(), res=cur)
            st["blocks"].append((cur, u, y16, hp, hh16, v, w49, w1_16t, w2_16t, h, w))
            cur = out
        tape["stages"].append(st)
    return cur, h, w, tape


def _trunk_bwd_tc(tr, tape, dcur, G):
    B = tape["B"]
    dims = tr.arch["dims"]
    for i in reversed(range(len(tr.stages))):
        stage, st, c = tr.stages[i], tape["stages"][i], dims[i]
        ones = torch.ones((c,), device=dcur.device, dtype=torch.float32)
        for blk, saved in zip(reversed(list(stage.blocks)), reversed(st["blocks"])):
            xin, u, y16, hp, hh16, v, w49, w1_16t, w2_16t, h, w = saved
            # dv = gamma * dcur: bf16 copies + its column sums (= d fc2.bias) + d gamma = sum(dcur * v) in ONE pass over dcur
            # the wgrad GEMMs read the same row-major bf16 copies as the forward / dgrad GEMMs (MN-major UMMA operands)
            dv16, _, db2, dgamma = cast_dual(dcur, t=False, colvec=blk.gamma.detach(), want_colsum=True, aux=v)
            G.put(blk.gamma, dgamma)
            G.put(blk.mlp.fc2.weight, tc_wgrad_mn(hh16, dv16).t().contiguous())         # [4c, c]^T -> [c, 4c]
            G.put(blk.mlp.fc2.bias, db2)
            dhh = tc_gemm16(dv16, w2_16t, torch.zeros((4 * c,), device=dcur.device, dtype=torch.float32))   # bf16 [M, 4c]
            dhp16, _, db1 = cast_dual(hp, t=False, op=2, x2=dhh, want_colsum=True)       # dhh * gelu'(hp)
            G.put(blk.mlp.fc1.weight, tc_wgrad_mn(dhp16, y16))                           # [4c, c]
            G.put(blk.mlp.fc1.bias, db1)
            dy = tc_gemm(dhp16, w1_16t)                                                  # [M, c]
            du, dlw, dlb = ln_bwd(u, blk.norm.weight.detach(), dy)
            G.put(blk.norm.weight, dlw)
            G.put(blk.norm.bias, dlb)
            dw49, dbias = dwconv_wgrad(xin, du, B, h, w)
            G.put(blk.conv_dw.weight, dw49.t().contiguous())
            G.put(blk.conv_dw.bias, dbias)
            dconv = dwconv(du, w49, None, B, h, w, flip=1)
            dcur = colscale(dconv, ones, res=dcur)
        if i > 0:
            xin, pt16, wds16t, h, w = st["down"]
            cin = dims[i - 1]
            ln_m, conv_m = stage.downsample[0], stage.downsample[1]
            d16, _, dbd = cast_dual(dcur, t=False, want_colsum=True)
            dwds = tc_wgrad_mn(pt16, d16).t().contiguous()          # [4cin, c]^T -> [c, (dy,dx,cin)]
            G.put(conv_m.weight, dwds.view(c, 2, 2, cin).permute(0, 3, 1, 2).contiguous())
            G.put(conv_m.bias, dbd)
            dpt = tc_gemm(d16, wds16t)                               # [Mo, 4cin]
            dyln = patch2x2(dpt, B, h, w, cin, reverse=True)
            dcur, dlw, dlb = ln_bwd(xin, ln_m.weight.detach(), dyln)
            G.put(ln_m.weight, dlw)
            G.put(ln_m.bias, dlb)
    patches, u0 = tape["stem"]
    stem_c, stem_n = tr.stem[0], tr.stem[1]
    du0, dlw, dlb = ln_bwd(u0, stem_n.weight.detach(), dcur)
    G.put(stem_n.weight, dlw)
    G.put(stem_n.bias, dlb)
    # stem wgrad [c0, 48] = du0^T patches on the tensor cores as well (the fp32 split-K kernel took 0.4 ms of the step)
    du16, _, db0 = cast_dual(du0, t=False, want_colsum=True)
    p16, _, _ = cast_dual(patches, t=False)
    G.put(stem_c.weight, tc_wgrad_mn(du16, p16))
    G.put(stem_c.bias, db0)


# ---- dense stacks (metadata branch, heads) -------------------------------------------------------------------------
class _Dense:
    """Sequential of Linear / act / Dropout / BatchNorm1d modules executed with the training kernels."""

    def __init__(self, modules, training=True, counter=None):
        self.mods = list(modules)
        self.training = training
        self.counter = counter
        self.saved = []

    def forward(self, x):
        import torch.nn as nn
        self.saved = []
        for m in self.mods:
            if isinstance(m, nn.Linear):
                self.saved.append(x)
                x = gemm_nt(x, m.weight.detach(), m.bias.detach())
            elif isinstance(m, (nn.GELU, nn.ReLU)):
                self.saved.append(x)
                x = act(x, L.ACT_GELU if isinstance(m, nn.GELU) else L.ACT_RELU)
            elif isinstance(m, nn.Dropout):
                if self.training and m.p > 0:
                    x, mask = dropout(x, m.p, seed=_seed(), counter=self.counter)
                    self.saved.append(mask)
                else:
                    self.saved.append(None)
            elif isinstance(m, nn.BatchNorm1d):
                B, F = x.shape
                y = torch.empty_like(x)
                mean, rstd = _new((F,), x), _new((F,), x)
                mom = BN_MOMENTUM if m.momentum is None else m.momentum
                L.launch("t_bn_fwd", L.lib().btsb_bn1d_train_fwd_f32, _p(x), _p(m.weight.detach()), _p(m.bias.detach()),
                         _p(m.running_mean), _p(m.running_var), float(mom), float(m.eps), _p(y), _p(mean), _p(rstd),
                         B, F, _st())
                m.num_batches_tracked += 1
                self.saved.append((x, mean, rstd))
                x = y
            else:
                raise TypeError(f"unsupported layer in dense stack: {type(m).__name__}")
        return x

    def backward(self, d, G, need_input_grad):
        import torch.nn as nn
        for idx in reversed(range(len(self.mods))):
            m, s = self.mods[idx], self.saved[idx]
            upstream = need_input_grad or any(_needs(*mm.parameters()) for mm in self.mods[:idx])
            if isinstance(m, nn.Linear):
                if m.weight.requires_grad:
                    G.put(m.weight, gemm_tn(d, s))
                    G.put(m.bias, colsum(d))
                d = gemm_nn(d, m.weight.detach()) if upstream else None
            elif isinstance(m, (nn.GELU, nn.ReLU)):
                d = act(s, L.ACT_GELU if isinstance(m, nn.GELU) else L.ACT_RELU, dout=d)
            elif isinstance(m, nn.Dropout):
                if s is not None:
                    d, _ = dropout(d, m.p, mask=s)
            elif isinstance(m, nn.BatchNorm1d):
                x, mean, rstd = s
                dx = torch.empty_like(x) if upstream else None
                dw, db = torch.zeros_like(m.weight), torch.zeros_like(m.bias)
                L.launch("t_bn_bwd", L.lib().btsb_bn1d_bwd_f32, _p(x), _p(d), _p(m.weight.detach()), _p(mean), _p(rstd),
                         _p(dx), _p(dw), _p(db), x.shape[0], x.shape[1], _st())
                G.put(m.weight, dw)
                G.put(m.bias, db)
                d = dx
            if d is None:
                return None
        return d


# ---- whole-model function ------------------------------------------------------------------------------------------
def _parts(model):
    """(trunk, pool_ln LayerNorm2d or None, meta modules or None, head modules or None) per reference class."""
    name = model._config["model_name"]
    if name == "mm_ConvNeXt":
        tr = model.convnext_backbone
        pl = tr.head[1] if isinstance(tr.head, torch.nn.Sequential) else None
        return tr, pl, list(model.metadata_branch), list(model.combined_head)
    if name == "ConvNeXt":
        tr = model.convnext
        return tr, tr.head[1], None, list(tr.head)[3:]
    if name == "um_nn":
        return None, None, list(model.network), None
    if name == "frozen_fusion":
        ib = model.image_branch
        if hasattr(ib, "convnext"):
            tr, pl = ib.convnext, ib.convnext.head[1]
        else:                                   # MaxViT image branch (architectures.py:304-308): frozen features only
            tr, pl = ib.maxvit, None
        return tr, pl, list(model.meta_branch.network), list(model.combined_head)
    raise NotImplementedError(
        f"btsbot_b200: no training (backward) path for model_name {name!r} -- MaxViT / mm_MaxViT run inference only "
        f"(call model.eval() and use torch.no_grad()); ConvNeXt, mm_ConvNeXt, um_nn and frozen_fusion train")


class _ModelFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, model, image, meta):
        L.lib()
        tr, pool_ln, meta_mods, head_mods = _parts(model)
        name = model._config["model_name"]
        ctx.model = model
        feat = None
        ctx.trunk_tape = None
        counter = getattr(model, "_graph_counter", None)        # set by GraphedTrainStep: device-side step counter
        if counter is not None:
            L.launch("t_counter", L.lib().btsb_counter_add_i64, _p(counter), 1, _st())
        if tr is not None:
            L.require_cuda(image, "image input")
            image = image.to(torch.float32).contiguous()
            B = image.shape[0]
            train_trunk = _needs(*tr.parameters())
            if train_trunk and type(tr).__name__ == "MaxVitTrunk":
                raise NotImplementedError("btsbot_b200: the MaxViT trunk has no backward path -- freeze the image branch "
                                          "(train.py:224-231 does for frozen_fusion) or use a ConvNeXt image branch")
            if train_trunk:
                tc = getattr(model, "_precision", "fp32") == "bf16"
                rows, h, w, tape = (_trunk_fwd_tc if tc else _trunk_fwd)(tr, image)
                ctx.trunk_tape = (tape, h, w)
                if pool_ln is not None:
                    pooled = pool(rows, B, h * w, rows.shape[1], reverse=False)
                    feat = ln_fwd(pooled, pool_ln.weight.detach(), pool_ln.bias.detach())
                    ctx.pool_saved = pooled
                else:
                    if h * w != 1:
                        raise RuntimeError("mat1 and mat2 shapes cannot be multiplied: the Flatten head needs a 1x1 map")
                    feat = rows
            else:
                # frozen image branch (frozen_fusion, train.py:224-231): features from the inference kernels
                feat = model._frozen_image_scorer().features(image).to(torch.float32)
        else:
            B = meta.shape[0]
        emb = None
        ctx.meta_stack = ctx.head_stack = None
        if meta_mods is not None:
            L.require_cuda(meta, "metadata input")
            ctx.meta_stack = _Dense(meta_mods, counter=counter)
            emb = ctx.meta_stack.forward(meta.to(torch.float32).contiguous())
        if head_mods is not None:
            cat = feat if emb is None else torch.cat((feat, emb), dim=1)
            ctx.split = feat.shape[1] if feat is not None else 0
            ctx.head_stack = _Dense(head_mods, counter=counter)
            logits = ctx.head_stack.forward(cat.contiguous())
        else:
            logits = emb                                            # um_nn: the stack ends in Linear(m2, 1)
        ctx.name = name
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        model = ctx.model
        tr, pool_ln, meta_mods, head_mods = _parts(model)
        G = _Grads(getattr(model, "_grad_sink", None))
        d = dlogits.to(torch.float32).contiguous()
        train_trunk = ctx.trunk_tape is not None
        meta_needs = meta_mods is not None and any(_needs(*m.parameters()) for m in meta_mods)
        dfeat = demb = None
        if ctx.head_stack is not None:
            dcat = ctx.head_stack.backward(d, G, need_input_grad=train_trunk or meta_needs)
            if dcat is not None:
                if meta_mods is not None:
                    dfeat = dcat[:, :ctx.split].contiguous() if train_trunk else None
                    demb = dcat[:, ctx.split:].contiguous() if meta_needs else None
                else:
                    dfeat = dcat
        else:
            demb = d
        if ctx.meta_stack is not None and demb is not None:
            ctx.meta_stack.backward(demb, G, need_input_grad=False)
        if train_trunk and dfeat is not None:
            tape, h, w = ctx.trunk_tape
            if pool_ln is not None:
                dpooled, dlw, dlb = ln_bwd(ctx.pool_saved, pool_ln.weight.detach(), dfeat)
                G.put(pool_ln.weight, dlw)
                G.put(pool_ln.bias, dlb)
                dfeat = pool(dpooled, tape["B"], h * w, dpooled.shape[1], reverse=True)
            (_trunk_bwd_tc if tape.get("tc") else _trunk_bwd)(tr, tape, dfeat, G)
        if G.sink is not None:
            G.sink.flush()
        return None, None, None, None


def training_forward(model, image_input=None, metadata_input=None):
    """Called by the model classes when ``model.training`` and grad mode is on (fp32 kernels; tensor-core GEMMs when
    the model's precision is ``"bf16"``)."""
    first = next((p for p in model.parameters() if p.requires_grad), None)
    if first is None:
        return model.scorer()(image_input=image_input, metadata_input=metadata_input)
    # The Function needs one input that requires grad so that its output does.  A fresh scalar leaf, made on the CURRENT
    # stream, instead of a parameter: a parameter's AccumulateGrad node is created once and remembers the stream of its
    # first use, and an eager step on the default stream followed by a CUDA-graph capture of the same model then makes the
    # autograd engine wait on that (uncaptured) stream -- cudaErrorStreamCaptureIsolation.
    anchor = torch.zeros((), device=first.device, dtype=torch.float32, requires_grad=True)
    return _ModelFn.apply(anchor, model, image_input, metadata_input)


class BCEWithLogitsLoss(torch.nn.Module):
    """``torch.nn.BCEWithLogitsLoss(pos_weight=...)`` (train.py:211-212) on one fused loss+gradient kernel."""

    def __init__(self, pos_weight=None):
        super().__init__()
        pw = 1.0 if pos_weight is None else float(torch.as_tensor(pos_weight).reshape(-1)[0])
        self.pos_weight_value = pw

    def forward(self, logits, labels):
        return _BCEFn.apply(logits, labels, self.pos_weight_value)


class _BCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, pw):
        L.require_cuda(logits, "logits")
        lg = logits.detach().to(torch.float32).contiguous()
        lb = labels.detach().to(device=lg.device, dtype=torch.float32).contiguous()
        if lg.numel() != lb.numel():
            raise ValueError(f"Target size ({tuple(labels.shape)}) must be the same as input size ({tuple(logits.shape)})")
        loss = torch.empty((1,), device=lg.device, dtype=torch.float32)
        dl = torch.empty_like(lg) if logits.requires_grad else None
        L.launch("t_bce", L.lib().btsb_bce_logits_f32, _p(lg), _p(lb), float(pw), _p(loss), _p(dl), lg.numel(), 1.0, _st())
        ctx.dl = dl
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        if ctx.dl is None:
            return None, None, None
        return ctx.dl * g, None, None


class FusedAdamW(torch.optim.Optimizer):
    """``torch.optim.AdamW(params, lr, betas)`` (train.py:242-246; weight_decay defaults to torch's 0.01) with one
    fused kernel per parameter tensor, or one per flat bucket when the parameters were flattened by
    ``btsbot_b200.parallel.flatten_parameters``."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, capturable=False):
        """``capturable=True`` (torch.optim's name for it) keeps the step count in a device tensor that the kernels read,
        so that ``step()`` can be captured in a CUDA graph and replayed (see :class:`GraphedTrainStep`)."""
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.capturable = bool(capturable)
        self._step_dev = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        lib = L.lib()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            by_step = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                L.require_cuda(p, "parameter")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["m"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["v"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                if not p.is_contiguous():
                    raise RuntimeError("FusedAdamW: parameters must be contiguous")
                by_step.setdefault(st["step"], []).append((p, p.grad.contiguous(), st))
            # parameters that share a step count (all of them, unless some had no gradient in earlier steps) go through
            # the multi-tensor kernel, BTSB_ADAMW_BATCH tensors per launch
            if self.capturable and by_step:
                dev = next(iter(by_step.values()))[0][0].device
                if self._step_dev is None:
                    self._step_dev = torch.zeros((1,), dtype=torch.int64, device=dev)
                L.launch("t_counter", lib.btsb_counter_add_i64, _p(self._step_dev), 1, _st())
            for step, items in by_step.items():
                for lo in range(0, len(items), L.ADAMW_BATCH):
                    part = items[lo:lo + L.ADAMW_BATCH]
                    batch = L.AdamwBatch()
                    for i, (p, g, st) in enumerate(part):
                        batch.p[i], batch.g[i], batch.m[i], batch.v[i] = p.data_ptr(), g.data_ptr(), st["m"].data_ptr(), st["v"].data_ptr()
                        batch.n[i] = p.numel()
                    batch.count = len(part)
                    if self.capturable:
                        L.launch("t_adamw", lib.btsb_adamw_multi_ctr_f32, C.byref(batch), float(group["lr"]), float(b1),
                                 float(b2), float(group["eps"]), float(group["weight_decay"]), _p(self._step_dev), 1.0, _st())
                    else:
                        L.launch("t_adamw", lib.btsb_adamw_multi_f32, C.byref(batch), float(group["lr"]), float(b1), float(b2),
                                 float(group["eps"]), float(group["weight_decay"]), int(step), 1.0, _st())
        # the kernels wrote the parameters through raw pointers: autograd's version counters did not move, so packed
        # inference copies (the models' Scorer cache) are invalidated explicitly
        from . import _engine
        _engine.bump_param_generation()
        return loss


class GraphedTrainStep:
    """One training step of ``train.py:496-547`` (zero_grad -> forward -> loss -> backward -> optimizer.step) captured once
    in a CUDA graph and replayed: the eager step issues ~360 kernels from Python (a third of them a few microseconds
    long), which makes the host the pacing side at batch 1024; a replay costs one launch.

    ``model`` is a btsbot_b200 model (or its DistributedDataParallel wrapper) in train mode, ``optimizer`` a
    :class:`FusedAdamW` built with ``capturable=True``, ``loss_fn`` a :class:`BCEWithLogitsLoss`.  Inputs of every call
    must have the shapes / dtypes of ``example``; they are copied into static buffers.  Dropout masks and the AdamW bias
    correction follow a device-side counter, so replays are not frozen at the captured step.  ``warmup`` eager steps run
    first (they DO update the parameters, like any other step).  With ``torch.distributed`` initialised and world size > 1
    the NCCL bucket all-reduces of the DistributedDataParallel wrapper are captured too (every rank captures and replays
    in lockstep; 1.34x faster than eager issue at N = 2); call :meth:`release` before destroying the process group."""

    def __init__(self, model, optimizer, loss_fn, example, warmup: int = 2):
        self.model, self.opt, self.loss_fn = model, optimizer, loss_fn
        if not getattr(optimizer, "capturable", False):
            raise ValueError("GraphedTrainStep needs FusedAdamW(..., capturable=True)")
        inner = getattr(model, "module", model)
        dev = next(inner.parameters()).device
        self.static = [t.to(dev).clone() if t is not None else None for t in example]
        inner._graph_counter = torch.zeros((1,), dtype=torch.int64, device=dev)
        import os
        import torch.distributed as dist
        # N > 1: the gradient all-reduces (NCCL, on the sink's side stream, forked from and joined back into the capture
        # stream) are captured with the rest of the step: 212 k alerts/s against 158 k eager at N = 2.  The graph must be
        # destroyed BEFORE the communicator (release()): tearing the process group down first hangs in NCCL.
        # BTSB_GRAPH_DDP=0 falls back to eager issue.
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.eager = multi and os.environ.get("BTSB_GRAPH_DDP", "1") == "0"
        self.graph, self.loss, self.kernels_per_step = None, None, 0
        if self.eager:
            return
        # warm-up steps on the capture side stream (lazy one-time initialisation -- kernel attributes, optimizer state,
        # autograd's per-stream bookkeeping -- must not happen during capture: capturing without one fails with
        # cudaErrorStreamCaptureIsolation).  They are real training steps on `example`; their loss / logits are kept in
        # `warmup_loss` / `warmup_logits` for callers that account for every batch (train_epoch).
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        self.warmup_loss = self.warmup_logits = None
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self.warmup_loss = self._step()
                self.warmup_logits = self.last_logits.clone()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = L.launch_count()
        with torch.cuda.graph(self.graph):
            self.loss = self._step()
        #: kernels of libbtsbot_b200.so inside one replay (the library's launch counter does not see replays)
        self.kernels_per_step = L.launch_count() - n0

    def _step(self):
        img, meta, lab = self.static
        self.model.zero_grad()
        if img is not None and meta is not None:
            logits = self.model(image_input=img, metadata_input=meta)
        else:
            logits = self.model(input_data=img if img is not None else meta)
        loss = self.loss_fn(logits, lab)
        loss.backward()
        self.opt.step()
        self.last_logits = logits.detach()          # static tensor inside the graph: overwritten by every replay
        return loss.detach()

    def release(self):
        """Destroy the captured graph (call before ``torch.distributed.destroy_process_group()`` when the NCCL all-reduce
        was captured: a communicator must outlive the graphs that use it)."""
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None

    def __call__(self, img, meta, lab):
        for dst, src in zip(self.static, (img, meta, lab)):
            if dst is not None:
                dst.copy_(src, non_blocking=True)
        if self.graph is None:
            return self._step()
        self.graph.replay()
        from . import _engine
        _engine.bump_param_generation()         # the replay re-ran the captured AdamW kernels
        return self.loss
