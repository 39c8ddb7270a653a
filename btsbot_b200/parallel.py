"""One-process-per-GPU data parallelism for the hot path.

* Inference / bulk scoring: alerts are independent, so rank r scores the contiguous index range
  ``shard_range(n, r, world)`` with **no collective on the data path**; :func:`score_alerts` optionally gathers the
  N float32 scores afterwards (host-side convenience, 4 B per alert).
* Training: replicas hold full weights; gradients are averaged with NCCL all-reduce on flat buckets that are launched
  from inside the hand-written backward (last layers first) on a side stream, overlapping the remaining backward.
  This replaces the reference's single-process ``nn.DataParallel`` (train.py:238-240), which re-broadcasts all
  parameters every step and reduces gradients to GPU 0.

Works with the ``gloo`` backend on CPU tensors too (used by the world-size-2 tests of the host logic).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def bind_to_device_numa(device_index: int):
    """Pin the calling process to the CPUs NVML reports as local to GPU ``device_index`` (its NUMA node), so that pinned
    host buffers allocated afterwards are first-touched next to the GPU's PCIe root: the H2D copy of a bulk-scoring
    step (391 MB per 8192 alerts) otherwise crosses the socket interconnect on multi-socket hosts and the end-to-end
    rate halves.  Returns the previous affinity set (restore it with ``os.sched_setaffinity(0, prev)``) or ``None``
    when NVML / affinity control is unavailable."""
    import os
    if os.environ.get("BTSB_NUMA_BIND", "1") == "0" or not hasattr(os, "sched_getaffinity"):
        return None
    try:
        import pynvml
        prev = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        try:
            bus = torch.cuda.get_device_properties(device_index).pci_bus_id
            for i in range(pynvml.nvmlDeviceGetCount()):
                h = pynvml.nvmlDeviceGetHandleByIndex(i)
                if int(pynvml.nvmlDeviceGetPciInfo(h).bus) == int(bus):
                    handle = h
        except Exception:
            pass
        n_words = (max(prev) // 64) + 1 if prev else 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, max(n_words, (os.cpu_count() or 64 + 63) // 64 + 1))
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        cpus &= prev
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return prev
    except Exception:
        return None


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced index range of ``rank``: sizes differ by at most one, concatenation order = rank order."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def score_alerts(score_fn, triplets, metadata, batch_size: int = 8192, gather: bool = True, device=None):
    """Score ``triplets [N,63,63,3]`` (+ ``metadata [N,M]``) sharded over the ranks of the default process group.

    ``score_fn(triplets_chunk, metadata_chunk) -> 1-D float32 tensor`` scores one micro-batch (normally
    :class:`AlertScorer`).  Returns the scores of the local shard (``gather=False``) or of all N alerts on every
    rank (``gather=True``; an all_gather of 4 bytes per alert after the data path)."""
    rank, world = _world()
    n = len(triplets)
    lo, hi = shard_range(n, rank, world)
    outs = []
    for a in range(lo, hi, batch_size):
        b = min(hi, a + batch_size)
        outs.append(score_fn(triplets[a:b], None if metadata is None else metadata[a:b]).reshape(-1).float())
    local = torch.cat(outs) if outs else torch.empty((0,), dtype=torch.float32, device=device)
    if not gather or world == 1:
        return local
    sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros((pad,), dtype=torch.float32, device=local.device)
    buf[: local.numel()] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)])


def _host_threads() -> int:
    """Host threads one rank may use for input marshalling: its share of the CPUs it is allowed to run on."""
    import os
    try:
        allowed = len(os.sched_getaffinity(0))
    except AttributeError:
        allowed = os.cpu_count() or 1
    local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    share = max(1, (os.cpu_count() or allowed) // local_world)
    return max(1, min(32, allowed, share))


class AlertScorer:
    """The end-to-end public scoring call: host (numpy / pinned torch) HWC triplets + metadata -> scores.

    Per micro-batch: async H2D copy on a copy stream into one of ``staging_slots`` device staging buffers the scorer owns,
    K1 (cast + NHWC->NCHW [+ crop/normalise]), model forward, sigmoid epilogue.  The copy of batch i+1 overlaps the kernels
    of batch i; a slot is overwritten only after the forward that read it has finished (one event per slot), so the host
    may run any number of calls ahead without the caching allocator ever entering the timed path (a fresh 391 MB
    ``.to(device)`` per call made the end-to-end rate swing between 0.65 and 1.13 M alerts/s from run to run:
    ``cudaMalloc`` in the middle of the pipeline whenever the host ran ahead of the events that free the old blocks).

    ``host_pack`` (bf16 ConvNeXt models, float32 triplets, plain cast + transpose preprocessing): the end-to-end rate of
    that path is the PCIe rate of the fp32 input (391 MB per 8192 alerts), and the first thing the bf16 trunk does with a
    pixel is round it to bf16 -- so the scorer can round a FRACTION f of each batch on the host (``btsb_host_pack_bf16``: a
    pool of host threads writing into pinned memory) while the DMA engine is already moving the other, untouched fp32
    part; the packed part follows with half its bytes, K1 runs once per part, and the logits are bit-identical.
    ``"auto"`` (default) times one pack and one fp32 copy of the first batch of each shape and picks the f that balances
    the host threads against the PCIe link (0 = plain copy when the host cannot keep up: few cores per GPU, small batches),
    then lets the pipeline decide: six calls plain, six calls split, the shorter measured period stays (single-rank jobs
    only: ranks sharing a host interfere in ways a per-rank measurement does not see);
    ``True`` / ``False`` / a float in [0, 1] force it, the environment variable ``BTSB_HOST_PACK`` (0, 1 or a fraction) too."""

    def __init__(self, model, crop_to_size: int = 63, normalize: bool = False, return_scores: bool = True,
                 staging_slots: int = 3, host_pack="auto"):
        import os
        from . import alert_utils, ops
        self.model, self.crop, self.norm, self.return_scores = model.eval(), crop_to_size, normalize, return_scores
        self._au, self._ops = alert_utils, ops
        self.multimodal = model._config["model_name"] in ("mm_ConvNeXt", "mm_MaxViT", "frozen_fusion")
        self.meta_only = model._config["model_name"] == "um_nn"
        self.dev = next(model.parameters()).device
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.n_slots = max(2, int(staging_slots))
        self._rings = {}                 # (shape, dtype) -> [next slot, [device buffer, consumed event | None] * n_slots]
        env = os.environ.get("BTSB_HOST_PACK")
        if env is not None:
            host_pack = float(env)
        if host_pack is True:
            host_pack = 1.0
        elif host_pack is False:
            host_pack = 0.0
        self.host_pack = host_pack       # "auto" or the packed fraction f
        self._pack_ok = int(crop_to_size) == 63 and not normalize and not self.meta_only and self._rounds_input(model)
        self._pack_rings = {}            # shape -> [next slot, [pinned bf16, dev bf16, dev f32, copied ev, consumed ev] * n]
        self._pack_choice = {}           # shape -> packed fraction f ("auto" calibration result)
        self.pack_threads = _host_threads()
        self._multi_rank = int(os.environ.get("LOCAL_WORLD_SIZE", "1")) > 1
        self.last_calibration = None     # (shape, pack_ms, copy_ms, f) of the latest "auto" decision
        self.last_fraction = 0.0         # packed fraction of the latest call
        self._probe = {}                 # shape -> state of the plain-vs-split A/B of the first calls ("auto" mode)
        self.last_probe = None           # {"plain_ms", "split_ms", "fraction", "kept"} of the latest decision

    @staticmethod
    def _rounds_input(model) -> bool:
        """True when the model's first use of a pixel is its bf16 rounding: the bf16 ConvNeXt trunks, whose stem is a
        tensor-core GEMM over bf16 patches (MaxViT interpolates the fp32 image first; the fp32 mode keeps fp32)."""
        from . import _engine, _lib as L
        try:
            tr = getattr(model.scorer(), "trunk", None)
        except Exception:
            return False
        return tr is not None and tr.code == L.BF16 and _engine.TC_STEM and tr.stem_w_tc is not None

    def _stage(self, x):
        """Queue the H2D copy of ``x`` into the next staging slot of its shape on the copy stream; returns the slot."""
        t = torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x.contiguous()
        if t.device == self.dev:
            return [t, None]
        key = (tuple(t.shape), t.dtype)
        ring = self._rings.get(key)
        if ring is None:
            if len(self._rings) >= 4:                        # a new batch shape: drop the oldest ring
                self._rings.pop(next(iter(self._rings)))
            ring = self._rings[key] = [0, [[torch.empty(t.shape, dtype=t.dtype, device=self.dev), None]
                                           for _ in range(self.n_slots)]]
        slot = ring[1][ring[0]]
        ring[0] = (ring[0] + 1) % self.n_slots
        if slot[1] is not None:
            self.copy_stream.wait_event(slot[1])             # the forward that read this slot last has finished
        slot[0].copy_(t, non_blocking=True)
        return slot

    # ---- split staging: rows [0, n1) rounded to bf16 on the host, rows [n1, B) copied as they are ----------------------
    def _pack(self, t, pinned):
        from . import _lib as L
        L.check(L.lib().btsb_host_pack_bf16(t.data_ptr(), pinned.data_ptr(), t.numel(), self.pack_threads), "host_pack")

    def _stage_split(self, t, f):
        B = t.shape[0]
        n1 = min(B, max(1, int(round(f * B))))
        key = tuple(t.shape)
        ring = self._pack_rings.get(key)
        if ring is None:
            if len(self._pack_rings) >= 2:
                self._pack_rings.pop(next(iter(self._pack_rings)))
            # full-size buffers, used through [:n1] / [n1:] views: the packed fraction may change from call to call
            ring = self._pack_rings[key] = [0, [[torch.empty(t.shape, dtype=torch.bfloat16).pin_memory(),
                                                 torch.empty(t.shape, dtype=torch.bfloat16, device=self.dev),
                                                 torch.empty(t.shape, dtype=torch.float32, device=self.dev),
                                                 None, None] for _ in range(self.n_slots)]]
        slot = ring[1][ring[0]]
        ring[0] = (ring[0] + 1) % self.n_slots
        if slot[4] is not None:
            self.copy_stream.wait_event(slot[4])             # the forward that read this slot's device buffers has finished
        if n1 < B:
            slot[2][n1:].copy_(t[n1:], non_blocking=True)    # the DMA engine starts on the fp32 part right away ...
        if slot[3] is not None:
            slot[3].synchronize()                            # (the copy that last read this pinned buffer has finished)
        self._pack(t[:n1], slot[0][:n1])                     # ... while the host threads round the other part
        slot[1][:n1].copy_(slot[0][:n1], non_blocking=True)
        if slot[3] is None:
            slot[3] = torch.cuda.Event()
        slot[3].record(self.copy_stream)
        self.last_fraction = n1 / float(B)
        return slot, n1

    def _packed_fraction(self, t) -> float:
        if not self._pack_ok or t.dtype != torch.float32 or t.device.type != "cpu" or t.dim() != 4 \
                or tuple(t.shape[1:]) != (63, 63, 3) or t.shape[0] == 0:
            return 0.0
        if self.host_pack != "auto":
            return float(self.host_pack)
        if self._multi_rank:
            # several ranks on one host share its memory system, and their packs and copies drift in and out of phase:
            # the per-rank probe read 5.0 ms (split) against 7.0 ms (plain) at N = 2 while the job's max-over-ranks step
            # was 7.8 ms against 7.3 ms plain (profiles/r02n2); N = 4 gained 9 % on one box, N = 8 with four threads per
            # rank lost.  "auto" therefore packs in single-rank jobs only; a fraction can still be forced.
            return 0.0
        key = tuple(t.shape)
        f = self._pack_choice.get(key)
        if f is None:
            f = self._pack_choice[key] = self._calibrate(t)
            if f > 0.0:
                self._probe[key] = {"f": f, "calls": 0, "events": [], "dt": ([], [])}
        if key in self._probe:                               # A/B of the first calls: plain, then split (see _probe_record)
            self._probe_harvest(key)
        pr = self._probe.get(key)
        if pr is not None:
            pr["calls"] += 1
            return 0.0 if pr["calls"] <= self.PROBE_CALLS else pr["f"]
        return self._pack_choice[key]

    PROBE_CALLS = 6

    def _probe_record(self, key, cur):
        """"auto" mode: the calibration predicts, the pipeline decides.  The first PROBE_CALLS calls of a batch shape take
        the plain copy, the next PROBE_CALLS the split; the time between the completions of consecutive forwards (CUDA
        events on the compute stream, harvested without blocking) is the pipeline period of either mode, measured with
        every rank of the job loading the host at once -- what a stand-alone calibration cannot see (8 ranks on a 32-vCPU
        box: predicted 6.0 ms, measured 20.7 ms per step with 30 % packed, profiles/r02n8).  The split stays only if its
        median period is at least 3 % shorter."""
        pr = self._probe.get(key)
        if pr is None:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(cur)
        pr["events"].append((ev, 0 if pr["calls"] <= self.PROBE_CALLS else 1))
        self._probe_harvest(key)

    def _probe_harvest(self, key):
        pr = self._probe[key]
        evs = pr["events"]
        while len(evs) >= 2 and evs[1][0].query():
            if evs[0][1] == evs[1][1]:
                pr["dt"][evs[1][1]].append(evs[0][0].elapsed_time(evs[1][0]))
            evs.pop(0)
        a, b = pr["dt"]
        if len(a) >= 4 and len(b) >= 4:
            ma, mb = sorted(a)[len(a) // 2], sorted(b)[len(b) // 2]
            keep = mb < 0.97 * ma
            self._pack_choice[key] = pr["f"] if keep else 0.0
            self.last_probe = {"plain_ms": ma, "split_ms": mb, "fraction": pr["f"], "kept": keep}
            del self._probe[key]

    def _calibrate(self, t) -> float:
        """One pack against one fp32 copy of this batch.  With a fraction f packed, a step occupies the calling thread
        for f * pack + the launches of the step, and the PCIe link for (1 - f/2) * copy; f balances the two, capped so that
        the pack is no longer than the link work queued beside it (else the link idles waiting for the pack).  An online
        hill-climb of f on the observed call period was tried on top of this (profiles/r02z/visit_log_adapt_*.txt): 1.51 M
        against 1.60 M alerts/s at N = 1 (the optimum is sharp, the search dithers around it) and 2.59 M against 2.60 M at
        N = 4 on a box whose host memory system, not f, limits the step -- removed."""
        import time
        # batches under 64 MB of fp32 (1400 alerts) keep the plain copy: their step is a millisecond, of which waking the
        # pack threads and the second K1 launch are a visible part (C2, 1024 alerts: 0.93 ms plain, 1.25 ms split)
        if t.numel() < (1 << 24):
            return 0.0
        probe = torch.empty(t.shape, dtype=torch.bfloat16).pin_memory()
        self._pack(t, probe)                                 # warm: wakes the pool, touches the pages
        pack_ms = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            self._pack(t, probe)
            pack_ms = min(pack_ms, (time.perf_counter() - t0) * 1e3)
        dst = torch.empty(t.shape, dtype=t.dtype, device=self.dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        src = t if t.is_pinned() else t.pin_memory()         # the plain path's best case
        with torch.cuda.stream(self.copy_stream):
            dst.copy_(src, non_blocking=True)
            e0.record(self.copy_stream)
            dst.copy_(src, non_blocking=True)
            e1.record(self.copy_stream)
        e1.synchronize()
        copy_ms = e0.elapsed_time(e1)
        # launches of a step on the calling thread: ~1.5 ms for the ~40 kernels of a forward issued from Python, ~0.3 ms
        # as one graph replay
        enqueue_ms = 0.3 if self.model._config.get("infer_cuda_graph") else 1.5
        # balance thread and link; keep the pack within the link work queued beside it (the previous batch's bf16 part +
        # this batch's fp32 part), and never beyond 0.65: measured on a 16-vCPU B200 box (profiles/r02z, 8192 alerts per
        # step, pack 4.9-6.9 ms, copy 7.1-7.3 ms) the step takes 7.25 / 6.13 / 5.56 / 5.15 / 5.75 / 5.96 / 6.34 ms at
        # f = 0 / 0.35 / 0.5 / 0.65 / 0.7 / 0.8 / 1.0
        f = (copy_ms - enqueue_ms) / (pack_ms + 0.5 * copy_ms)
        f = max(0.0, min(f, copy_ms / (pack_ms + 0.5 * copy_ms), 0.65))
        period = max(f * pack_ms + enqueue_ms, (1.0 - 0.5 * f) * copy_ms)
        if period > 0.9 * copy_ms:                           # not worth the host threads
            f = 0.0
        self.last_calibration = (tuple(t.shape), pack_ms, copy_ms, f)
        return f

    @torch.no_grad()
    def __call__(self, triplets, metadata=None):
        cur = torch.cuda.current_stream(self.dev)
        ts = ps = None
        n1 = 0
        if not self.meta_only:
            th = torch.from_numpy(np.ascontiguousarray(triplets)) if isinstance(triplets, np.ndarray) else triplets
            f = self._packed_fraction(th.contiguous()) if th.device.type == "cpu" else 0.0
            with torch.cuda.stream(self.copy_stream):
                if f > 0.0:
                    ps, n1 = self._stage_split(th.contiguous(), f)
                else:
                    ts = self._stage(th)
                    self.last_fraction = 0.0
        with torch.cuda.stream(self.copy_stream):
            ms = self._stage(metadata) if (self.multimodal or self.meta_only) else None
        cur.wait_stream(self.copy_stream)
        m = None if ms is None else ms[0]
        if self.meta_only:
            logits = self.model(input_data=m)
        else:
            if ps is not None:                               # K1 once per part, into one [B,3,63,63] tensor
                B = ps[2].shape[0]
                x = torch.empty((B, 3, 63, 63), device=self.dev, dtype=torch.float32)
                self._au.triplets_to_model_input(ps[1][:n1], self.crop, self.norm, out=x[:n1])
                if n1 < B:
                    self._au.triplets_to_model_input(ps[2][n1:], self.crop, self.norm, out=x[n1:])
            else:
                x = self._au.triplets_to_model_input(ts[0], self.crop, self.norm)
            logits = self.model(image_input=x, metadata_input=m) if self.multimodal else self.model(input_data=x)
        for slot in (ts, ms):
            if slot is not None:
                if slot[1] is None:
                    slot[1] = torch.cuda.Event()
                slot[1].record(cur)
        if ps is not None:
            if ps[4] is None:
                ps[4] = torch.cuda.Event()
            ps[4].record(cur)
        if self._probe and not self.meta_only:
            self._probe_record(tuple(th.shape), cur)
        if not self.return_scores:
            return logits.reshape(-1)
        scores, _ = self._ops.score(logits)
        return scores.reshape(-1)


# ---------------------------------------------------------------------------------------------------------------
# training: flat gradient buckets + overlapped all-reduce
# ---------------------------------------------------------------------------------------------------------------
class GradSink:
    """Flat fp32 gradient buffer; parameters are laid out in REVERSE forward order so the backward fills it front to
    back, and every completed bucket is all-reduced (average) immediately on a side stream."""

    def __init__(self, params, bucket_bytes: int = 8 << 20, process_group=None):
        self.params = [p for p in params if p.requires_grad]
        self.pg = process_group
        order = list(reversed(self.params))
        self.offset, off = {}, 0
        for p in order:
            self.offset[id(p)] = off
            off += p.numel()
        self.total = off
        dev = self.params[0].device
        self.flat = torch.zeros((self.total,), device=dev, dtype=torch.float32)
        per = max(1, bucket_bytes // 4)
        self.bounds = [(a, min(self.total, a + per)) for a in range(0, self.total, per)]
        self.bucket_of = {}
        for p in order:
            a = self.offset[id(p)]
            b = a + p.numel()
            self.bucket_of[id(p)] = [i for i, (lo, hi) in enumerate(self.bounds) if lo < b and a < hi]
        self.need = [0] * len(self.bounds)
        for p in order:
            for i in self.bucket_of[id(p)]:
                self.need[i] += 1
        self.cuda = dev.type == "cuda"
        self.stream = torch.cuda.Stream(device=dev) if self.cuda else None
        self.last_order, self.last_early = [], 0     # bucket launch order / #buckets sent before flush() (last step)
        #: False skips the collectives (bench.py times a compute-only step to derive the all-reduce overlap fraction)
        self.comm = True
        self.reset()

    def reset(self):
        self.count = [0] * len(self.bounds)
        self.seen = set()
        self.sent = [False] * len(self.bounds)
        self.handles = []
        self.launched_buckets = []
        self.in_flush = False
        self.early = 0
        self.side_used = False          # a collective was issued on the side stream in this step

    def view(self, p):
        a = self.offset[id(p)]
        return self.flat[a:a + p.numel()].view(p.shape)

    def adopt(self, p, g):
        v = self.view(p)
        v.copy_(g)
        return v

    def ready(self, p):
        if id(p) in self.seen:
            return
        self.seen.add(id(p))
        for i in self.bucket_of[id(p)]:
            self.count[i] += 1
            if self.count[i] == self.need[i]:
                self._launch(i)

    def _launch(self, i):
        if self.sent[i]:
            return
        self.sent[i] = True
        self.launched_buckets.append(i)
        if not self.in_flush:
            self.early += 1
        _, world = _world()
        if world == 1 or not self.comm:
            return
        lo, hi = self.bounds[i]
        chunk = self.flat[lo:hi]
        if self.cuda:
            self.side_used = True
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                dist.all_reduce(chunk, op=dist.ReduceOp.AVG, group=self.pg)
        else:
            self.handles.append((dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.pg, async_op=True), chunk, world))

    def flush(self):
        """End of backward: send what is left (parameters without gradient this step), then join."""
        self.in_flush = True
        for i in range(len(self.bounds)):
            self._launch(i)
        if self.cuda and self.side_used:                  # join the side stream only if this step forked work into it (a
            torch.cuda.current_stream().wait_stream(self.stream)   # capturing stream must not wait on uncaptured work)
        for h, chunk, world in self.handles:
            h.wait()
            chunk.div_(world)
        self.last_order, self.last_early = self.launched_buckets, self.early
        self.reset()


class DistributedDataParallel(torch.nn.Module):
    """``model = DistributedDataParallel(model)``: same role as ``DataParallel(model)`` at train.py:238-240, but one
    process per GPU.  ``.module`` is the wrapped model (the reference unwraps it when saving, train.py:314-317)."""

    def __init__(self, module, bucket_mb: float = 8.0, process_group=None):
        super().__init__()
        self.module = module
        rank, world = _world()
        if world > 1:
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=0, group=process_group)
        self.sink = GradSink(list(module.parameters()), int(bucket_mb * (1 << 20)), process_group)
        module._grad_sink = self.sink

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def zero_grad(self, set_to_none: bool = True):
        self.module.zero_grad(set_to_none=set_to_none)
