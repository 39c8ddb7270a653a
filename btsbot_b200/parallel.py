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


class AlertScorer:
    """The end-to-end public scoring call: host (numpy / pinned torch) HWC triplets + metadata -> scores.

    Per micro-batch: async H2D copy on a copy stream into one of ``staging_slots`` device staging buffers the scorer owns,
    K1 (cast + NHWC->NCHW [+ crop/normalise]), model forward, sigmoid epilogue.  The copy of batch i+1 overlaps the kernels
    of batch i; a slot is overwritten only after the forward that read it has finished (one event per slot), so the host
    may run any number of calls ahead without the caching allocator ever entering the timed path (a fresh 391 MB
    ``.to(device)`` per call made the end-to-end rate swing between 0.65 and 1.13 M alerts/s from run to run:
    ``cudaMalloc`` in the middle of the pipeline whenever the host ran ahead of the events that free the old blocks)."""

    def __init__(self, model, crop_to_size: int = 63, normalize: bool = False, return_scores: bool = True,
                 staging_slots: int = 3):
        from . import alert_utils, ops
        self.model, self.crop, self.norm, self.return_scores = model.eval(), crop_to_size, normalize, return_scores
        self._au, self._ops = alert_utils, ops
        self.multimodal = model._config["model_name"] in ("mm_ConvNeXt", "mm_MaxViT", "frozen_fusion")
        self.meta_only = model._config["model_name"] == "um_nn"
        self.dev = next(model.parameters()).device
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.n_slots = max(2, int(staging_slots))
        self._rings = {}                 # (shape, dtype) -> [next slot, [device buffer, consumed event | None] * n_slots]

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

    @torch.no_grad()
    def __call__(self, triplets, metadata=None):
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.copy_stream):
            ts = None if self.meta_only else self._stage(triplets)
            ms = self._stage(metadata) if (self.multimodal or self.meta_only) else None
        cur.wait_stream(self.copy_stream)
        t, m = (None if ts is None else ts[0]), (None if ms is None else ms[0])
        if self.meta_only:
            logits = self.model(input_data=m)
        else:
            x = self._au.triplets_to_model_input(t, self.crop, self.norm)
            logits = self.model(image_input=x, metadata_input=m) if self.multimodal else self.model(input_data=x)
        for slot in (ts, ms):
            if slot is not None:
                if slot[1] is None:
                    slot[1] = torch.cuda.Event()
                slot[1].record(cur)
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
