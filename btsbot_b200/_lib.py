"""ctypes binding of ``libbtsbot_b200.so`` (the C ABI declared in ``include/btsbot_b200.h``).

There is deliberately no fallback: if the shared library is missing or a call fails, a
``RuntimeError`` is raised.  The product never computes on the CPU or through PyTorch ops.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

F32, BF16, F64, BF16_XF16 = 0, 1, 2, 3      # BF16_XF16: bf16 compute, the residual-stream tensor of the call in fp16
ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
EPI_BIAS, EPI_BIAS_GELU, EPI_SCALE_RES, EPI_BIAS_SILU = 0, 1, 2, 3

_HERE = os.path.dirname(os.path.abspath(__file__))
#: the in-tree build; BTSB_LIB names another build of the same sources (A/B timing of a compile-time variant) -- still
#: this library or nothing, there is no fallback
LIB_PATH = os.environ.get("BTSB_LIB") or os.path.join(_HERE, "libbtsbot_b200.so")

_lib = None
_lock = threading.Lock()

vp, i32, i64, f32p = C.c_void_p, C.c_int, C.c_int64, C.c_void_p


class HeadParams(C.Structure):
    """``btsb_head_params`` (include/btsbot_b200.h)."""
    _fields_ = [
        ("feat", vp), ("feat_dtype", i32), ("F", i32),
        ("meta", vp), ("Mm", i32),
        ("bn_scale", vp), ("bn_shift", vp),
        ("m1t", vp), ("m1b", vp), ("m1", i32),
        ("m2t", vp), ("m2b", vp), ("m2", i32),
        ("meta_act", i32), ("meta_out_act", i32),
        ("h0t", vp), ("h0b", vp), ("c1", i32),
        ("h1t", vp), ("h1b", vp), ("c2", i32),
        ("h2", vp), ("h2b", vp),
        ("head_act", i32),
        ("h0_init", vp),
    ]


ADAMW_BATCH = 48


class AdamwBatch(C.Structure):
    """``btsb_adamw_batch`` (include/btsbot_b200.h)."""
    _fields_ = [("p", vp * ADAMW_BATCH), ("g", vp * ADAMW_BATCH), ("m", vp * ADAMW_BATCH), ("v", vp * ADAMW_BATCH),
                ("n", i64 * ADAMW_BATCH), ("first_block", i32 * (ADAMW_BATCH + 1)), ("count", i32)]


#: every symbol include/btsbot_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "btsb_version": (i32, []),
    "btsb_last_error_string": (C.c_char_p, []),
    "btsb_device_ok": (i32, []),
    "btsb_launch_count": (C.c_uint64, []),
    "btsb_preprocess_crop_norm": (i32, [vp, i32, i64, i32, i32, i32, vp, vp]),
    "btsb_host_pack_bf16": (i32, [vp, vp, i64, i32]),
    "btsb_preprocess_pad_norm": (i32, [vp, vp, i64, i32, vp, i32, vp, vp]),
    "btsb_ingest_fits_gz": (i32, [vp, vp, i64, vp, vp, i32]),
    "btsb_augment_gather_f32": (i32, [vp, vp, vp, i64, i32, vp, vp]),
    "btsb_convnext_stem_fwd": (i32, [vp, i64, i32, i32, vp, vp, vp, vp, i32, vp, i32, vp]),
    "btsb_convnext_dwln_fwd": (i32, [vp, i32, i64, i32, i32, i32, vp, vp, vp, vp, vp, vp]),
    "btsb_convnext_lnpatch_fwd": (i32, [vp, i32, i64, i32, i32, i32, vp, vp, vp, vp]),
    "btsb_convnext_poolln_fwd": (i32, [vp, i32, i64, i32, i32, vp, vp, vp, vp]),
    "btsb_gemm_fwd": (i32, [vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, vp]),
    "btsb_stem_im2col_bf16": (i32, [vp, vp, i64, i32, i32, vp]),
    "btsb_stem_fused_fwd": (i32, [vp, i64, i32, i32, vp, vp, vp, vp, vp, i32, i32, vp]),
    "btsb_gemm_ln_fwd": (i32, [vp, vp, vp, vp, vp, vp, i64, i32, i32, vp]),
    "btsb_convnext_mlp_fused_fwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, vp]),
    "btsb_meta_head_fwd": (i32, [C.POINTER(HeadParams), i64, vp, vp]),
    "btsb_score_epilogue": (i32, [vp, i64, vp, vp, vp]),
    "btsb_gemm_f32_strided": (i32, [vp, i64, i64, vp, i64, i64, vp, i64, i64, i64, i32, vp]),
    "btsb_colsum_f32": (i32, [vp, vp, vp, i64, i32, i32, vp]),
    "btsb_act_f32": (i32, [vp, vp, vp, i64, i32, vp]),
    "btsb_colscale_f32": (i32, [vp, vp, vp, vp, i64, i32, vp]),
    "btsb_bias_add_f32": (i32, [vp, vp, i64, i32, vp]),
    "btsb_layernorm_fwd_f32": (i32, [vp, vp, vp, vp, i64, i32, C.c_float, vp]),
    "btsb_layernorm_bwd_f32": (i32, [vp, vp, vp, vp, vp, vp, i64, i32, C.c_float, vp]),
    "btsb_dwconv7_f32": (i32, [vp, vp, vp, vp, i64, i32, i32, i32, i32, vp]),
    "btsb_dwconv7_wgrad_f32": (i32, [vp, vp, vp, vp, i64, i32, i32, i32, vp]),
    "btsb_stem_im2col_f32": (i32, [vp, vp, i64, i32, i32, vp]),
    "btsb_patch2x2_f32": (i32, [vp, vp, i64, i32, i32, i32, i32, vp]),
    "btsb_pool_f32": (i32, [vp, vp, i64, i32, i32, i32, vp]),
    "btsb_bn1d_train_fwd_f32": (i32, [vp, vp, vp, vp, vp, C.c_float, C.c_float, vp, vp, vp, i64, i32, vp]),
    "btsb_bn1d_bwd_f32": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, vp]),
    "btsb_dropout_f32": (i32, [vp, vp, vp, i64, C.c_float, C.c_uint64, i32, vp]),
    "btsb_bce_logits_f32": (i32, [vp, vp, C.c_float, vp, vp, i64, C.c_float, vp]),
    "btsb_adamw_f32": (i32, [vp, vp, vp, vp, i64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, i64,
                             C.c_float, vp]),
    "btsb_adamw_multi_f32": (i32, [C.POINTER(AdamwBatch), C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, i64,
                                   C.c_float, vp]),
    "btsb_adamw_multi_ctr_f32": (i32, [C.POINTER(AdamwBatch), C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, vp,
                                       C.c_float, vp]),
    "btsb_dropout_ctr_f32": (i32, [vp, vp, vp, i64, C.c_float, C.c_uint64, vp, i32, vp]),
    "btsb_counter_add_i64": (i32, [vp, i64, vp]),
    "btsb_maxvit_stem1_fwd": (i32, [vp, i64, i32, i32, i32, vp, vp, i32, vp, i32, vp]),
    "btsb_maxvit_im2col3_fwd": (i32, [vp, vp, i64, i32, i32, i32, i32, vp]),
    "btsb_conv3x3_c32_fwd": (i32, [vp, vp, vp, vp, i64, i32, i32, i32, vp]),
    "btsb_maxvit_avgpool2_fwd": (i32, [vp, vp, i64, i32, i32, i32, i32, vp]),
    "btsb_maxvit_dw3_fwd": (i32, [vp, i64, i32, i32, i32, i32, vp, vp, vp, vp, i32, vp]),
    "btsb_maxvit_se_fwd": (i32, [vp, i64, i32, i32, vp, vp, vp, vp, vp, vp]),
    "btsb_maxvit_scale_fwd": (i32, [vp, vp, i64, i32, i32, i32, vp]),
    "btsb_layernorm_rows_fwd": (i32, [vp, vp, vp, vp, i64, i32, i32, vp]),
    "btsb_maxvit_attn_fwd": (i32, [vp, vp, i64, i32, i32, i32, i32, vp, i32, vp]),
    "btsb_maxvit_lnpool_fwd": (i32, [vp, vp, vp, vp, i64, i32, i32, i32, vp]),
    "btsb_debug_mlp_trace": (i32, [vp]),
    "btsb_cast_dual_bf16": (i32, [vp, vp, vp, vp, vp, vp, i64, i32, i64, i32, i32, vp, vp, vp]),
    "btsb_gemm_bf16_f32out": (i32, [vp, vp, vp, vp, i64, i32, i32, vp]),
    "btsb_gemm_bf16_wgrad": (i32, [vp, vp, i64, vp, i32, i32, i64, vp]),
    "btsb_gemm_bf16_wgrad_mn": (i32, [vp, vp, vp, i32, i32, i64, vp]),
    "btsb_cast_f32_to_bf16": (i32, [vp, vp, i64, vp]),
    "btsb_cast_bf16_to_f32": (i32, [vp, vp, i64, vp]),
}


def lib() -> C.CDLL:
    """Load (once) and return the CUDA library; raises if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.isfile(LIB_PATH):
                    raise RuntimeError(
                        f"btsbot_b200: CUDA extension {LIB_PATH} is missing -- build it with "
                        f"`python -c 'import __graft_entry__ as g; g.build()'` or btsbot_b200/csrc/build.sh. "
                        f"There is no CPU fallback.")
                handle = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)     # AttributeError if the .so is stale
                    fn.restype, fn.argtypes = res, args
                _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().btsb_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"btsbot_b200 {what} failed (code {rc}): {msg}")


class KernelProfiler:
    """Optional per-launch CUDA-event timing (bench.py's roofline leg).  Events are recorded on the stream the
    kernel is launched on (torch's current stream); ``summary()`` synchronises and aggregates by kernel name."""

    def __init__(self):
        self.records = []          # (name, ev0, ev1, flops, bytes)

    def summary(self):
        import torch
        torch.cuda.synchronize()
        agg = {}
        for name, e0, e1, fl, by in self.records:
            a = agg.setdefault(name, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
            a["launches"] += 1
            a["ms"] += e0.elapsed_time(e1)
            a["flops"] += fl
            a["bytes"] += by
        return agg


#: set to a KernelProfiler to time every launch made through :func:`launch`
profiler = None


def launch(name: str, fn, *args, flops: float = 0.0, nbytes: float = 0.0) -> None:
    """Call one C-ABI kernel entry point, raise on failure, optionally bracket it with CUDA events.
    ``flops`` / ``nbytes`` are the ALGORITHMIC work of this launch (DESIGN.md "roofline accounting")."""
    prof = profiler
    if prof is None:
        check(fn(*args), name)
        return
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(fn(*args), name)
    e1.record()
    prof.records.append((name, e0, e1, flops, nbytes))


def launch_count() -> int:
    return int(lib().btsb_launch_count())


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t, name: str):
    if not t.is_cuda:
        raise RuntimeError(
            f"btsbot_b200: `{name}` is on {t.device}; this package only runs on an sm_100a GPU "
            f"(no CPU fallback) -- move the model and its inputs to 'cuda'.")
