"""Host-side driver of the sm_100a kernels: weight packing and the kernel sequence of one forward pass.

PyTorch is used here only for device memory (``torch.empty``), streams and one-off weight re-layout at
load time; every arithmetic step of the hot path is a call into ``libbtsbot_b200.so`` through ctypes
(``_lib``).  Nothing in this file can run without the CUDA library and an sm_100 device.

Layout contract: API tensors are NCHW float32 like the reference's; between
kernels activations are NHWC pixel rows ``[B*H*W, C]`` in the compute dtype (float32 or bfloat16).
"""
from __future__ import annotations

import os

import ctypes as C

import torch

from . import _lib as L
from .synth import convnext_arch

BN_EPS = 1e-5  # torch.nn.BatchNorm1d default, used by the reference (architectures.py:147)

#: one-kernel stem (im2col rows built in shared memory by producer warps); BTSB_STEM_FUSED=0 -> im2col + GEMM-with-LN
STEM_FUSED = os.environ.get("BTSB_STEM_FUSED", "1") != "0"
#: the wide variants of the fused kernel (C = 256 / 320: single D2 accumulator, G2 split in two UMMAs); BTSB_FUSE_WIDE=0
#: falls back to the two separate GEMMs for A/B timing
FUSE_MLP_WIDE = os.environ.get("BTSB_FUSE_WIDE", "1") != "0"
#: wide fused MLP called IN PLACE (out == res): the update is added to the residual stream by a bulk tensor reduction
#: instead of load -> add -> store: 126 -> 117 us per launch at C = 320 (profiles/r02t); BTSB_MLP_INPLACE=0 for A/B
MLP_INPLACE = os.environ.get("BTSB_MLP_INPLACE", "1") != "0"
#: bf16 mode: the residual stream of every stage but the last (stem / downsample outputs, block outputs) is stored as
#: IEEE fp16 instead of bf16 -- same bytes, 8x finer rounding of the tensor that is updated 12-14 times in a row; the
#: MMA operands stay bf16, the last stage stays bf16 because its rows feed the head GEMM (csrc/common.cuh).
#: BTSB_XF16=0 restores the all-bf16 stream for A/B.
XF16 = os.environ.get("BTSB_XF16", "1") != "0"
#: use the fused fc1->GELU->fc2 kernel where it applies (bf16, C <= 160, 256, 320); tests flip this to cover both paths
FUSE_MLP = True
#: head layer 0: contract the F image features on the tensor cores (bf16 features x bf16 weights, fp32 accumulate and
#: output, bias included) and let the head kernel add the metadata-embedding rows: meta_head 137 -> 50 us + 15 us GEMM
#: per 8192 alerts, logits within 1.2e-3 of the all-fp32 head; BTSB_HEAD_TC=0 keeps everything in the head kernel
HEAD_TC = os.environ.get("BTSB_HEAD_TC", "1") != "0"
#: bf16 stem as im2col + tcgen05 GEMM with the LayerNorm in the epilogue (else the CUDA-core stem kernel)
TC_STEM = True

_DT = {"fp32": (L.F32, torch.float32), "bf16": (L.BF16, torch.bfloat16)}

#: bumped by every in-place parameter update that bypasses autograd's version counters (FusedAdamW.step writes through
#: raw device pointers; a CUDA-graph replay re-runs those kernels) -- part of the models' Scorer cache key
_PARAM_GENERATION = 0


def param_generation() -> int:
    return _PARAM_GENERATION


def bump_param_generation() -> None:
    global _PARAM_GENERATION
    _PARAM_GENERATION += 1


#: kernels of libbtsbot_b200.so executed through CUDA-graph replays of the inference forward (the library's own launch
#: counter only sees launches issued from the host, i.e. the capture pass)
_GRAPH_REPLAY_LAUNCHES = 0


def graph_replay_launches() -> int:
    return _GRAPH_REPLAY_LAUNCHES


class _GraphedForward:
    """One inference forward of a :class:`Scorer` at a fixed input shape, captured in a CUDA graph (config key
    ``infer_cuda_graph``).  At small batches (BASELINE config C2: 1024 alerts, C1: 39) the ~40-kernel forward is paced by
    the host: every launch encodes its tensor maps and crosses ctypes, and the GPU idles between kernels.  A replay is one
    launch; the tensor maps are encoded once, at capture.  Inputs are copied into static buffers, the logits are returned
    as a fresh tensor."""

    def __init__(self, scorer, image, meta):
        self.img = image.clone() if image is not None else None
        self.meta = meta.clone() if meta is not None else None
        dev = (image if image is not None else meta).device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                       # lazy one-time initialisation must not happen during capture
            for _ in range(2):
                scorer(image_input=self.img, metadata_input=self.meta)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        n0 = L.launch_count()
        with torch.cuda.graph(self.graph):
            self.out = scorer(image_input=self.img, metadata_input=self.meta)
        self.kernels = L.launch_count() - n0

    def __call__(self, image, meta):
        global _GRAPH_REPLAY_LAUNCHES
        if self.img is not None:
            self.img.copy_(image, non_blocking=True)
        if self.meta is not None:
            self.meta.copy_(meta, non_blocking=True)
        self.graph.replay()
        _GRAPH_REPLAY_LAUNCHES += self.kernels
        return self.out.clone()


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


class TrunkWeights:
    """Kernel-ready copy of a timm-keyed ConvNeXt trunk (keys: SURVEY.md section 8b)."""

    def __init__(self, sd: dict, prefix: str, arch: dict, precision: str):
        code, wdt = _DT[precision]
        self.precision, self.code, self.wdt = precision, code, wdt
        self.dims, self.depths = tuple(arch["dims"]), tuple(arch["depths"])
        g = lambda k: sd[prefix + k]
        c0 = self.dims[0]
        # stem.0 Conv2d weight [C0,3,4,4] -> [48, C0] with k = (ci*4+ky)*4+kx
        self.stem_w = _f32c(g("stem.0.weight").reshape(c0, 48).t())
        self.stem_b = _f32c(g("stem.0.bias"))
        self.stem_ln_w, self.stem_ln_b = _f32c(g("stem.1.weight")), _f32c(g("stem.1.bias"))
        # tensor-core stem (bf16): [C0, 64] with the 48 taps zero-padded to one 64-wide k-block
        self.stem_w_tc = None
        if code == L.BF16 and c0 % 16 == 0 and c0 <= 128:
            wt = torch.zeros((c0, 64), device=self.stem_b.device, dtype=torch.bfloat16)
            wt[:, :48] = g("stem.0.weight").detach().reshape(c0, 48).to(torch.bfloat16)
            self.stem_w_tc = wt.contiguous()
        self.stages = []
        for i, (c, d) in enumerate(zip(self.dims, self.depths)):
            st = {"blocks": []}
            if i > 0:
                cin = self.dims[i - 1]
                q = f"stages.{i}.downsample."
                st["ds_ln_w"], st["ds_ln_b"] = _f32c(g(q + "0.weight")), _f32c(g(q + "0.bias"))
                # Conv2d [Cout,Cin,2,2] -> GEMM weight [Cout, (dy,dx,cin)] matching the lnpatch column order
                st["ds_w"] = g(q + "1.weight").detach().permute(0, 2, 3, 1).reshape(c, 4 * cin).to(wdt).contiguous()
                st["ds_b"] = _f32c(g(q + "1.bias"))
            for j in range(d):
                q = f"stages.{i}.blocks.{j}."
                st["blocks"].append(dict(
                    dw_w=_f32c(g(q + "conv_dw.weight").reshape(c, 49).t()),      # [49, C], k = ky*7+kx
                    dw_b=_f32c(g(q + "conv_dw.bias")),
                    ln_w=_f32c(g(q + "norm.weight")), ln_b=_f32c(g(q + "norm.bias")),
                    fc1_w=g(q + "mlp.fc1.weight").detach().reshape(4 * c, c).to(wdt).contiguous(),
                    fc1_b=_f32c(g(q + "mlp.fc1.bias")),
                    fc2_w=g(q + "mlp.fc2.weight").detach().reshape(c, 4 * c).to(wdt).contiguous(),
                    fc2_b=_f32c(g(q + "mlp.fc2.bias")),
                    gamma=_f32c(g(q + "gamma")),
                ))
            self.stages.append(st)


def _gemm(name, a, wt, bias, gamma, res, out, code, epi, st):
    """out = epi(a @ wt^T + bias); algorithmic work: 2MNK flops; bytes = A + W + out (+ residual)."""
    M, K = a.shape
    N = wt.shape[0]
    es = a.element_size()
    nbytes = es * (M * K + N * K + M * N + (M * N if res is not None else 0)) + 4.0 * N
    L.launch(name, L.lib().btsb_gemm_fwd, _p(a), _p(wt), _p(bias), _p(gamma), _p(res), _p(out), M, N, K, code, epi, st,
             flops=2.0 * M * N * K, nbytes=nbytes)


def trunk_forward(w: TrunkWeights, x: torch.Tensor, capture: dict | None = None):
    """timm ``forward_features`` on the GPU kernels.  ``x``: [B,3,H,W] float32 CUDA.
    Returns ``(rows [B*h*w, C_last] in the compute dtype, h, w)``."""
    lib = L.lib()
    L.require_cuda(x, "image input")
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError(f"expected image batch [B,3,H,W], got {tuple(x.shape)}")
    x = x.to(torch.float32).contiguous()
    B, _, H, W = x.shape
    if H < 4 or W < 4:
        raise ValueError("image smaller than the 4x4 patch stem")
    dev, st = x.device, L.stream_ptr()
    code, adt = w.code, w.wdt
    h, wd = (H - 4) // 4 + 1, (W - 4) // 4 + 1
    c = w.dims[0]
    last = len(w.stages) - 1
    stem_fused = TC_STEM and STEM_FUSED and w.stem_w_tc is not None and c in (64, 80, 96)
    # fp16 residual stream (bf16 mode, every stage but the last): needs the kernels that know the BF16_XF16 code
    xf16 = XF16 and code == L.BF16 and stem_fused and last >= 1

    def stream(i):
        """(dtype code, torch dtype) of stage i's residual stream."""
        return (L.BF16_XF16, torch.float16) if xf16 and i < last else (code, adt)

    scode, sdt = stream(0)
    cur = torch.empty((B * h * wd, c), device=dev, dtype=sdt)
    es = cur.element_size()
    if stem_fused:
        L.launch("stem_fused", lib.btsb_stem_fused_fwd, _p(x), B, H, W, _p(w.stem_w_tc), _p(w.stem_b), _p(w.stem_ln_w),
                 _p(w.stem_ln_b), _p(cur), c, scode, st, flops=2.0 * 48 * c * B * h * wd,
                 nbytes=4.0 * x.numel() + es * cur.numel())
    elif TC_STEM and w.stem_w_tc is not None:
        patches = torch.empty((B * h * wd, 64), device=dev, dtype=torch.bfloat16)
        L.launch("stem_im2col", lib.btsb_stem_im2col_bf16, _p(x), _p(patches), B, H, W, st,
                 nbytes=4.0 * x.numel() + 2.0 * patches.numel())
        L.launch("stem_gemm_ln", lib.btsb_gemm_ln_fwd, _p(patches), _p(w.stem_w_tc), _p(w.stem_b), _p(w.stem_ln_w),
                 _p(w.stem_ln_b), _p(cur), B * h * wd, c, 64, st, flops=2.0 * 48 * c * B * h * wd,
                 nbytes=2.0 * patches.numel() + es * cur.numel())
    else:
      L.launch("stem", lib.btsb_convnext_stem_fwd, _p(x), B, H, W, _p(w.stem_w), _p(w.stem_b), _p(w.stem_ln_w),
               _p(w.stem_ln_b), c, _p(cur), code, st,
               flops=2.0 * 48 * c * B * h * wd, nbytes=4.0 * x.numel() + es * cur.numel())
    if capture is not None:
        capture["stem"] = (cur, h, wd)
    for i, stg in enumerate(w.stages):
        c = w.dims[i]
        if i > 0:
            cin = w.dims[i - 1]
            if h < 2 or wd < 2:
                raise ValueError("feature map too small for the 2x2/s2 downsample")
            ho, wo = (h - 2) // 2 + 1, (wd - 2) // 2 + 1
            patches = torch.empty((B * ho * wo, 4 * cin), device=dev, dtype=adt)
            L.launch("lnpatch", lib.btsb_convnext_lnpatch_fwd, _p(cur), scode, B, h, wd, cin, _p(stg["ds_ln_w"]),
                     _p(stg["ds_ln_b"]), _p(patches), st, flops=8.0 * patches.numel(),
                     nbytes=es * (cur.numel() + patches.numel()))
            h, wd = ho, wo
            scode, sdt = stream(i)
            cur = torch.empty((B * h * wd, c), device=dev, dtype=sdt)
            _gemm("gemm_down", patches, stg["ds_w"], stg["ds_b"], None, None, cur, scode, L.EPI_BIAS, st)
            if capture is not None:
                capture[f"down{i}"] = (cur, h, wd)
        M = B * h * wd
        y = torch.empty((M, c), device=dev, dtype=adt)
        fused = FUSE_MLP and code == L.BF16 and c % 16 == 0 and (64 <= c <= 160 or (FUSE_MLP_WIDE and c in (256, 320)))
        hid = None if fused else torch.empty((M, 4 * c), device=dev, dtype=adt)
        for j, blk in enumerate(stg["blocks"]):
            L.launch(f"dwln_{wd}x{c}", lib.btsb_convnext_dwln_fwd, _p(cur), scode, B, h, wd, c, _p(blk["dw_w"]),
                     _p(blk["dw_b"]), _p(blk["ln_w"]), _p(blk["ln_b"]), _p(y), st,
                     flops=2.0 * 49 * M * c + 8.0 * M * c, nbytes=2.0 * es * M * c)
            if capture is not None:
                capture[f"s{i}b{j}.dwln"] = (y.clone(), h, wd)
            # the block updates its input rows in place where the kernel can add to them (wide fused MLP: bulk tensor
            # reduction); `capture` keeps every block's output, so it takes the out-of-place form
            inplace = fused and MLP_INPLACE and c in (256, 320) and capture is None
            nxt = cur if inplace else torch.empty((M, c), device=dev, dtype=sdt)
            if fused:
                # fc1 -> GELU -> fc2 -> *gamma -> +shortcut in one kernel; bytes: y + res + out (+ L2-resident weights)
                L.launch(f"mlp_fused_{c}", lib.btsb_convnext_mlp_fused_fwd, _p(y), _p(cur), _p(blk["fc1_w"]),
                         _p(blk["fc1_b"]), _p(blk["fc2_w"]), _p(blk["fc2_b"]), _p(blk["gamma"]), _p(nxt), M, c,
                         L.BF16_XF16 if scode == L.BF16_XF16 else L.BF16, st,
                         flops=16.0 * M * c * c, nbytes=es * (3.0 * M * c + 8.0 * c * c))
            else:
                _gemm(f"gemm_fc1_{c}", y, blk["fc1_w"], blk["fc1_b"], None, None, hid, code, L.EPI_BIAS_GELU, st)
                _gemm(f"gemm_fc2_{c}", hid, blk["fc2_w"], blk["fc2_b"], blk["gamma"], cur, nxt, scode,
                      L.EPI_SCALE_RES, st)
            cur = nxt
            if capture is not None:
                capture[f"s{i}b{j}"] = (cur, h, wd)
    return cur, h, wd


def pool_ln(rows: torch.Tensor, B: int, hw: int, ln_w, ln_b, code: int) -> torch.Tensor:
    """global-avg-pool + LayerNorm2d + flatten -> [B,C] float32 (architectures.py:109-113,136-141)."""
    lib = L.lib()
    cdim = rows.shape[1]
    out = torch.empty((B, cdim), device=rows.device, dtype=torch.float32)
    L.launch("poolln", lib.btsb_convnext_poolln_fwd, _p(rows), code, B, hw, cdim, _p(ln_w), _p(ln_b), _p(out),
             L.stream_ptr(), flops=8.0 * rows.numel(), nbytes=rows.element_size() * rows.numel() + 4.0 * out.numel())
    return out


class HeadWeights:
    """Folded BatchNorm1d + transposed fp32 Linear weights for the fused metadata/head kernel."""

    def __init__(self, sd: dict, *, meta_prefix=None, head_prefix=None, head_idx=(0, 2, 5), final_prefix=None,
                 meta_act=L.ACT_GELU, meta_out_act=L.ACT_GELU, head_act=L.ACT_GELU):
        t = lambda k: _f32c(sd[k].t())
        self.meta_act, self.meta_out_act, self.head_act = meta_act, meta_out_act, head_act
        self.Mm = self.m1 = self.m2 = self.c1 = self.c2 = 0
        self.bn_scale = self.bn_shift = self.m1t = self.m1b = self.m2t = self.m2b = None
        self.h0t = self.h0b = self.h1t = self.h1b = None
        if meta_prefix is not None:
            p = meta_prefix
            var, mean = sd[p + "0.running_var"].float(), sd[p + "0.running_mean"].float()
            scale = sd[p + "0.weight"].float() / torch.sqrt(var + BN_EPS)
            self.bn_scale = scale.contiguous()
            self.bn_shift = (sd[p + "0.bias"].float() - mean * scale).contiguous()
            self.m1t, self.m1b = t(p + "1.weight"), _f32c(sd[p + "1.bias"])
            self.m2t, self.m2b = t(p + "4.weight"), _f32c(sd[p + "4.bias"])
            self.Mm, self.m1, self.m2 = self.m1t.shape[0], self.m1t.shape[1], self.m2t.shape[1]
        if head_prefix is not None:
            a, b, c = head_idx
            self.h0t, self.h0b = t(f"{head_prefix}{a}.weight"), _f32c(sd[f"{head_prefix}{a}.bias"])
            self.h1t, self.h1b = t(f"{head_prefix}{b}.weight"), _f32c(sd[f"{head_prefix}{b}.bias"])
            self.c1, self.c2 = self.h0t.shape[1], self.h1t.shape[1]
            self._h0w = sd[f"{head_prefix}{a}.weight"]          # [c1, F + m2]; bf16 feature slice made on first use
            self._h0w16 = {}
            final_prefix = f"{head_prefix}{c}."
        self.h2 = _f32c(sd[final_prefix + "weight"].reshape(-1))
        self.h2b = _f32c(sd[final_prefix + "bias"].reshape(-1))
        self.in_features = self.h0t.shape[0] if self.h0t is not None else self.m2


def _head_feature_gemm(hw: HeadWeights, feat: torch.Tensor) -> torch.Tensor:
    """``h0b + feat @ h0w[:, :F]^T`` as fp32 ``[B, c1]`` on tcgen05 (the 82 k of the head's 120 k MACs per alert)."""
    B, F = feat.shape
    w16 = hw._h0w16.get(F)
    if w16 is None:
        w16 = hw._h0w16[F] = hw._h0w.detach()[:, :F].to(torch.bfloat16).contiguous()
    out = torch.empty((B, hw.c1), device=feat.device, dtype=torch.float32)
    L.launch("head_gemm", L.lib().btsb_gemm_bf16_f32out, _p(feat), _p(w16), _p(hw.h0b), _p(out), B, hw.c1, F,
             L.stream_ptr(), flops=2.0 * B * hw.c1 * F, nbytes=2.0 * (B * F + hw.c1 * F) + 4.0 * B * hw.c1)
    return out


def head_forward(hw: HeadWeights, feat: torch.Tensor | None, meta: torch.Tensor | None, B: int) -> torch.Tensor:
    lib = L.lib()
    p = L.HeadParams()
    dev = (feat if feat is not None else meta).device
    F = 0
    h0_init = None
    if feat is not None:
        feat = feat.contiguous()
        F = feat.shape[1]
        p.feat, p.F = feat.data_ptr(), F
        p.feat_dtype = L.BF16 if feat.dtype == torch.bfloat16 else L.F32
        if (HEAD_TC and feat.dtype == torch.bfloat16 and 0 < hw.c1 <= 2560 and hw.c1 % 16 == 0 and F % 8 == 0
                and F + (hw.m2 if meta is not None else 0) == hw.in_features):
            h0_init = _head_feature_gemm(hw, feat)
            p.h0_init = h0_init.data_ptr()
    if meta is not None:
        L.require_cuda(meta, "metadata input")
        meta = meta.to(torch.float32).contiguous()
        if meta.dim() != 2 or meta.shape[1] != hw.Mm or meta.shape[0] != B:
            raise ValueError(f"expected metadata [{B},{hw.Mm}], got {tuple(meta.shape)}")
        p.meta, p.Mm = meta.data_ptr(), hw.Mm
        p.bn_scale, p.bn_shift = hw.bn_scale.data_ptr(), hw.bn_shift.data_ptr()
        p.m1t, p.m1b, p.m1 = hw.m1t.data_ptr(), hw.m1b.data_ptr(), hw.m1
        p.m2t, p.m2b, p.m2 = hw.m2t.data_ptr(), hw.m2b.data_ptr(), hw.m2
    p.meta_act, p.meta_out_act, p.head_act = hw.meta_act, hw.meta_out_act, hw.head_act
    if hw.c1 > 0:
        if F + (hw.m2 if meta is not None else 0) != hw.in_features:
            # the reference fails the same way inside nn.Linear when the trunk map is not 1x1 (architectures.py:142-143)
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied: head expects {hw.in_features} "
                               f"input features, got {F + (hw.m2 if meta is not None else 0)}")
        p.h0t, p.h0b, p.c1 = hw.h0t.data_ptr(), hw.h0b.data_ptr(), hw.c1
        p.h1t, p.h1b, p.c2 = hw.h1t.data_ptr(), hw.h1b.data_ptr(), hw.c2
    p.h2, p.h2b = hw.h2.data_ptr(), hw.h2b.data_ptr()
    logits = torch.empty((B, 1), device=dev, dtype=torch.float32)
    macs = hw.Mm * hw.m1 + hw.m1 * hw.m2 + hw.in_features * hw.c1 + hw.c1 * hw.c2 + max(hw.c2, 1)
    nbytes = B * (F * (feat.element_size() if feat is not None else 0) + 4.0 * hw.Mm + 4.0)
    if h0_init is not None:                                 # the feature rows went through head_gemm
        macs -= F * hw.c1
        nbytes = B * (4.0 * hw.c1 + 4.0 * hw.Mm + 4.0)
    L.launch("meta_head", lib.btsb_meta_head_fwd, C.byref(p), B, _p(logits), L.stream_ptr(), flops=2.0 * macs * B,
             nbytes=nbytes)
    return logits


class Scorer:
    """Eval-mode forward of one reference model class on the B200 kernels.

    ``sd`` is a reference/timm-keyed state dict whose tensors live on the target CUDA device (e.g.
    ``module.state_dict()``).  Mirrors `btsbot/architectures.py` ``forward`` of mm_ConvNeXt (:166-171),
    ConvNeXt (:121-122), um_nn (:292-293), frozen_fusion (:366-372), MaxViT (:42-51) and mm_MaxViT (:88-101).
    """

    def __init__(self, config: dict, sd: dict, precision: str = "fp32"):
        if precision not in _DT:
            raise ValueError(f"precision must be 'fp32' or 'bf16', got {precision!r}")
        L.lib()
        self.config, self.precision = config, precision
        self.name = name = config["model_name"]
        self.trunk = self.head = self.maxvit = None
        self.pool_ln = None
        self._graphs = {}                                   # input shapes -> _GraphedForward (config infer_cuda_graph)
        if name == "mm_ConvNeXt":
            arch = convnext_arch(config.get("model_kind", "convnext_nano.d1h_in1k"))
            self.trunk = TrunkWeights(sd, "convnext_backbone.", arch, precision)
            if "LS" in config["train_data_version"]:
                self.pool_ln = (_f32c(sd["convnext_backbone.head.1.weight"]), _f32c(sd["convnext_backbone.head.1.bias"]))
            self.head = HeadWeights(sd, meta_prefix="metadata_branch.", head_prefix="combined_head.")
        elif name == "ConvNeXt":
            arch = convnext_arch(config.get("model_kind", "convnext_nano.d1h_in1k"))
            self.trunk = TrunkWeights(sd, "convnext.", arch, precision)
            self.pool_ln = (_f32c(sd["convnext.head.1.weight"]), _f32c(sd["convnext.head.1.bias"]))
            self.head = HeadWeights(sd, head_prefix="convnext.head.", head_idx=(3, 5, 8))
        elif name == "mm_MaxViT":
            from . import _maxvit
            self.maxvit = _maxvit.MaxVitWeights(sd, "maxvit_backbone.", _maxvit.arch_for(config), precision)
            self.head = HeadWeights(sd, meta_prefix="metadata_branch.", head_prefix="combined_head.")
        elif name == "MaxViT":
            from . import _maxvit
            self.maxvit = _maxvit.MaxVitWeights(sd, "maxvit.", _maxvit.arch_for(config), precision)
            self.head = HeadWeights(sd, head_prefix="maxvit.head.", head_idx=(1, 3, 6))
        elif name == "um_nn":
            self.head = HeadWeights(sd, meta_prefix="network.", final_prefix="network.6.",
                                    meta_act=L.ACT_RELU, meta_out_act=L.ACT_RELU)
        elif name == "frozen_fusion":
            icfg = config["image_model_config"]
            if config["meta_model_config"]["model_name"] != "um_nn":
                raise ValueError("B200 frozen_fusion path: the metadata branch must be um_nn")
            if icfg["model_name"] == "ConvNeXt":            # head cut to [pool, LayerNorm2d, flatten] (architectures.py:309-313)
                arch = convnext_arch(icfg.get("model_kind", "convnext_nano.d1h_in1k"))
                self.trunk = TrunkWeights(sd, "image_branch.convnext.", arch, precision)
                self.pool_ln = (_f32c(sd["image_branch.convnext.head.1.weight"]),
                                _f32c(sd["image_branch.convnext.head.1.bias"]))
            elif icfg["model_name"] == "MaxViT":            # head cut to [pool] (architectures.py:304-308): what the
                from . import _maxvit                       # published maxvit-tiny-*-metadata checkpoints are (to_HF.py:143)
                self.maxvit = _maxvit.MaxVitWeights(sd, "image_branch.maxvit.", _maxvit.arch_for(icfg), precision)
            else:
                raise ValueError(f"B200 frozen_fusion path: image branch {icfg['model_name']!r} is not supported "
                                 f"(ConvNeXt or MaxViT)")
            self.head = HeadWeights(sd, meta_prefix="meta_branch.network.", head_prefix="combined_head.",
                                    meta_act=L.ACT_RELU, meta_out_act=L.ACT_NONE, head_act=L.ACT_RELU)
        else:
            raise ValueError(f"no B200 path for model_name {name!r}")

    def features(self, image_input: torch.Tensor, capture: dict | None = None) -> torch.Tensor:
        if self.maxvit is not None:
            from . import _maxvit
            return _maxvit.maxvit_features(self.maxvit, image_input, capture)
        rows, h, w = trunk_forward(self.trunk, image_input, capture)
        B = image_input.shape[0]
        if self.pool_ln is not None:
            return pool_ln(rows, B, h * w, self.pool_ln[0], self.pool_ln[1], self.trunk.code)
        if h * w != 1:
            # nn.Flatten(1) of [B,C,h,w] is channel-major; only reachable when the reference itself would
            # fail in combined_head (in_features == C), so just build the tensor it would have built
            return rows.view(B, h * w, -1).permute(0, 2, 1).reshape(B, -1).contiguous()
        return rows

    def graphed(self, image_input=None, metadata_input=None) -> torch.Tensor:
        """``__call__`` through a per-input-shape CUDA graph (at most 8 shapes are kept)."""
        uses_img = self.trunk is not None or self.maxvit is not None
        img = meta = None
        if uses_img:
            L.require_cuda(image_input, "image input")
            img = image_input.to(torch.float32).contiguous()
            if img.shape[0] == 0:
                return self(image_input=image_input, metadata_input=metadata_input)
        if self.head.Mm > 0:
            L.require_cuda(metadata_input, "metadata input")
            meta = metadata_input.to(torch.float32).contiguous()
            if meta.shape[0] == 0:
                return self(image_input=image_input, metadata_input=metadata_input)
        key = (tuple(img.shape) if img is not None else None, tuple(meta.shape) if meta is not None else None)
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= 8:
                self._graphs.pop(next(iter(self._graphs)))
            self(image_input=img, metadata_input=meta)      # argument errors surface eagerly, not inside a capture
            g = self._graphs[key] = _GraphedForward(self, img, meta)
        return g(img, meta)

    def __call__(self, image_input=None, metadata_input=None, capture: dict | None = None) -> torch.Tensor:
        feat = None
        if self.trunk is not None or self.maxvit is not None:
            B = image_input.shape[0]
            if B > 0:
                feat = self.features(image_input, capture)
        else:
            L.require_cuda(metadata_input, "metadata input")
            B = metadata_input.shape[0]
        meta = metadata_input if self.head.Mm > 0 else None
        if B == 0:
            dev = (image_input if image_input is not None else metadata_input).device
            return torch.empty((0, 1), device=dev, dtype=torch.float32)
        if capture is not None and feat is not None:
            capture["features"] = feat
        return head_forward(self.head, feat, meta, B)
