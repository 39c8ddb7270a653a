"""Array preparation of ZTF alert cutouts on the GPU -- drop-in names from `btsbot/alert_utils.py`.

* :func:`crop_norm_cutout`, :func:`crop_triplets`  (alert_utils.py:54-107) -> kernel ``btsb_preprocess_crop_norm``
* :func:`make_triplet` / :func:`make_triplets`     (alert_utils.py:110-196) -> batched multi-threaded gunzip + FITS parse
  in the library (``btsb_ingest_fits_gz``, no astropy, no Python per stamp), then kernel ``btsb_preprocess_pad_norm``
  for the numeric tail (nan_to_num, L2 normalise, drop flags, pad with 1e-9)
* :func:`triplets_to_model_input`  -- the cast + NHWC->NCHW step every reference caller does next
  (inference_example.py:62-64, train.py:139-155, val.py:92-94), fused with crop/normalise, output left on the GPU.
* :func:`extract_triplets` (alert_utils.py:199-226) host bookkeeping.

All arithmetic runs in ``libbtsbot_b200.so``; there is no numpy fallback.  Differences from the reference that
are deliberate: inputs are not mutated (the reference normalises the caller's array in place through a view,
alert_utils.py:75-76), and values are rounded to float32 once (callers cast to float32 anyway).
Plotting / Kowalski query helpers of the reference module are outside the hot path and not provided.
"""
from __future__ import annotations

import ctypes as C
import numpy as np
import torch

from . import _lib as L


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("btsbot_b200.alert_utils needs an sm_100a GPU (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _p(t):
    return C.c_void_p(t.data_ptr())


def triplets_to_model_input(triplets, crop_to_size: int = 63, normalize: bool = False, out=None) -> torch.Tensor:
    """``[N,63,63,3]`` HWC float32/float64 (numpy or torch, host or device) -> ``[N,3,s,s]`` float32 CUDA tensor.

    ``normalize=False`` is exactly ``astype(float32)`` + ``transpose(0,3,1,2)``; ``normalize=True`` additionally
    applies ``crop_triplets`` semantics (centre crop with margin ``(63-s)//2`` and per-cutout L2 normalisation).
    ``out``: an existing contiguous ``[N,3,s,s]`` float32 CUDA tensor (e.g. a slice of a batch) to write into."""
    return _crop(triplets, crop_to_size, normalize, out_hwc=False, out=out)


def _crop(triplets, s, normalize, out_hwc, out=None):
    lib = L.lib()
    dev = _dev()
    t = torch.from_numpy(np.ascontiguousarray(triplets)) if isinstance(triplets, np.ndarray) else triplets
    if t.dim() != 4 or tuple(t.shape[1:]) != (63, 63, 3):
        raise ValueError(f"expected triplets of shape [N,63,63,3], got {tuple(t.shape)}")
    # bfloat16 rows = triplets packed on the host for the scoring path (parallel.AlertScorer / btsb_host_pack_bf16):
    # cast + transpose only
    packed = t.dtype == torch.bfloat16 and s == 63 and not normalize and not out_hwc
    if t.dtype not in (torch.float32, torch.float64) and not packed:
        t = t.to(torch.float64)
    t = t.to(dev, non_blocking=True).contiguous()
    n = t.shape[0]
    shape = (n, s, s, 3) if out_hwc else (n, 3, s, s)
    if out is None:
        out = torch.empty(shape, device=dev, dtype=torch.float32)
    elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != t.device:
        raise ValueError(f"out must be a contiguous float32 tensor of shape {shape} on {t.device}")
    if n == 0:
        return out
    code = L.BF16 if packed else (L.F32 if t.dtype == torch.float32 else L.F64)
    # algorithmic bytes: read 63*63*3*e_in, write 3*s*s*4 per alert
    L.launch("crop_norm", lib.btsb_preprocess_crop_norm, _p(t), code, n, int(s), int(bool(normalize)), int(bool(out_hwc)),
             _p(out), L.stream_ptr(), flops=3.0 * n * 63 * 63 * 3,
             nbytes=float(n) * (63 * 63 * 3 * t.element_size() + 3 * s * s * 4))
    return out


def crop_triplets(triplets, crop_to_size):
    """Crop every cutout to ``crop_to_size`` and re-normalise with the L2 norm (alert_utils.py:81-107).
    Returns ``ndarray[N,s,s,3]`` float64 like the reference (values carry float32 precision)."""
    out = _crop(np.asarray(triplets), int(crop_to_size), True, out_hwc=True)
    return out.cpu().numpy().astype(np.float64)


def crop_norm_cutout(cutout, crop_to_size):
    """Single 63x63 cutout version (alert_utils.py:54-78)."""
    c = np.asarray(cutout)
    if c.shape != (63, 63):
        raise ValueError(f"expected a 63x63 cutout, got {c.shape}")
    trip = np.repeat(c[None, :, :, None], 3, axis=3)
    return crop_triplets(trip, crop_to_size)[0, :, :, 0].astype(c.dtype if c.dtype.kind == "f" else np.float64)


# ---- alert ingest: gz-FITS stamps -> triplets -------------------------------------------------------------
def decode_stamps(blobs, threads: int = 0):
    """Batched gunzip + FITS parse of alert stamps on a pool of host threads (``btsb_ingest_fits_gz``; replaces the
    per-stamp ``gzip.open`` + ``astropy.io.fits.open`` of alert_utils.py:141-147).  ``blobs``: sequence of ``bytes``.
    Returns ``(stamps float32 [n, 63*63] -- each stamp dense at the start of its slot --, hw int32 [n, 2])``."""
    n = len(blobs)
    keep = [bytes(b) for b in blobs]                           # own the buffers for the duration of the call
    ptrs = (C.c_char_p * n)(*keep)
    sizes = np.array([len(b) for b in keep], dtype=np.int64)
    stamps = np.zeros((n, 63 * 63), dtype=np.float32)
    hw = np.zeros((n, 2), dtype=np.int32)
    if n:
        L.check(L.lib().btsb_ingest_fits_gz(C.cast(ptrs, C.c_void_p), C.c_void_p(sizes.ctypes.data), n,
                                            C.c_void_p(stamps.ctypes.data), C.c_void_p(hw.ctypes.data), int(threads)),
                "ingest_fits_gz")
    return stamps, hw


def make_triplets(alerts, normalize: bool = True):
    """Batched :func:`make_triplet`: ``(ndarray[N,63,63,3] float64, ndarray[N] bool)``; one host call that inflates and
    parses all 3N stamps, one kernel launch for the numeric tail."""
    lib = L.lib()
    dev = _dev()
    n = len(alerts)
    blobs = [alert[f"cutout{which}"]["stampData"] for alert in alerts for which in ("Science", "Template", "Difference")]
    stamps, hw = decode_stamps(blobs)
    stamps, hw = stamps.reshape(n, 3, 63 * 63), hw.reshape(n, 3, 2)
    s_d = torch.from_numpy(stamps).to(dev)
    hw_d = torch.from_numpy(hw).to(dev)
    out = torch.empty((n, 63, 63, 3), device=dev, dtype=torch.float64)
    drop = torch.empty((n,), device=dev, dtype=torch.uint8)
    # algorithmic bytes: the staged float32 stamps in, the float64 63x63x3 triplet + drop flag out
    L.launch("pad_norm", lib.btsb_preprocess_pad_norm, _p(s_d), _p(hw_d), n, int(bool(normalize)), _p(out), L.F64, _p(drop),
             L.stream_ptr(), flops=3.0 * n * 63 * 63 * 3, nbytes=float(n) * (3 * 63 * 63 * 4 + 24 + 63 * 63 * 3 * 8 + 1))
    return out.cpu().numpy(), drop.cpu().numpy().astype(bool)


def make_triplet(alert, normalize: bool = True):
    """Unpack the three gzipped FITS stamps of one alert and pre-process them (alert_utils.py:110-196).
    Returns ``(triplet[63,63,3] float64, drop)``."""
    trip, drop = make_triplets([alert], normalize)
    return trip[0], bool(drop[0])


def extract_triplets(alerts):
    """Separate ``alert['triplet']`` from the alert dicts (alert_utils.py:199-226)."""
    triplets = np.empty((len(alerts), 63, 63, 3))
    for i, alert in enumerate(alerts):
        triplets[i] = alert["triplet"]
        alert.pop("triplet")
        alert.pop("cutoutScience")
        alert.pop("cutoutTemplate")
        alert.pop("cutoutDifference")
    return alerts, triplets
