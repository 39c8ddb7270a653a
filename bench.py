#!/usr/bin/env python
"""bench.py -- alerts/sec of the multimodal ConvNeXt scoring hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--precision bf16|fp32]
                    [--workload c3|c4|c5]

Workload (BASELINE.json configs[2], "C3"): bulk scoring of synthetic alerts with mm_ConvNeXt / convnext_nano
(random-init weights, 25 metadata columns), index-range sharded over ranks with NO collective on the data path.
A "step" is one forward pass of every rank over one micro-batch of B alerts ([B,3,63,63] fp32 + [B,25] fp32,
390 MB at B=8192 -> larger than the 126 MB L2, and consecutive steps rotate over distinct resident batches).

  value    : alerts/s, inputs already resident in HBM, CUDA events around exactly K steps, max over ranks
  e2e      : same metric through the public API (parallel.AlertScorer) from pinned HOST buffers: H2D copy of the HWC
             triplets + metadata (copy stream, overlapping the previous micro-batch's kernels), K1 layout kernel, model
             forward, D2H of the logits -- all inside the timed region
  roofline : dominant kernel (largest share of the step) timed per launch with CUDA events in an instrumented
             replay of the same K steps right after the timed region (events perturb launches, so `value`
             comes from the un-instrumented pass; `kernels` lists every kernel family for cross-checking)
  cpu_baseline : the CPU oracle (port of the reference path: btsbot/architectures.py glue + restated timm trunk)
             following inference_example.py:62-91 (batch 64, fp32, eval/no_grad) on a bounded sample

`--workload c4` runs BASELINE.json configs[3] instead (multimodal MaxViT-tiny-rw-224, batch 4096 per GPU per step, bf16);
`--workload c5` runs configs[4]: one TRAINING step (train.py:496-547: zero_grad, forward, BCE-with-logits, backward,
AdamW) of the multimodal ConvNeXt-nano on 1024 alerts per GPU in mixed precision (tcgen05 bf16 GEMMs); with one process
the whole step is ONE CUDA-graph replay (--no-graph issues it eagerly); with N > 1 the gradients are all-reduced over NCCL
on a side stream, overlapped with the backward, and those collectives are part of the captured graph.  The default (what the driver measures) is C3.

`--impl reference` times that CPU port alone with all host threads (the reference itself cannot run offline:
timm is not installable).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL_KIND = "convnext_nano.d1h_in1k"
ALERT_IN_BYTES = 63 * 63 * 3 * 4 + 25 * 4
#: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed `ncu --set full` captures at
#: 8192 alerts per launch (profiles/r01m/*.summary.txt; gemm_* from profiles/r01f)
NCU_TRAFFIC_8192 = {
    "mlp_fused_320": 96.09e6 + 13.36e6,
    "mlp_fused_80": 590.0e6 + 267.1e6,
    "dwln_15x80": 295.1e6 + 242.3e6,
    "gemm_fc1_320": 48.0e6 + 132.8e6,
    "gemm_fc2_320": 236.8e6 + 33.1e6,
}

WORKLOADS = {
    "c3": dict(model="mm_ConvNeXt", kind="convnext_nano.d1h_in1k", batch=8192, cpu_sample=8192, cpu_batch=64,
               label="C3 multimodal ConvNeXt-nano bulk scoring, 63x63x3 triplet + 25 metadata per alert"),
    "c4": dict(model="mm_MaxViT", kind="maxvit_tiny_rw_224.sw_in1k", batch=4096, cpu_sample=64, cpu_batch=64,
               label="C4 multimodal MaxViT-tiny-rw-224 scoring (bilinear 63->224, MBConv, window+grid attention), "
                     "63x63x3 triplet + 25 metadata per alert"),
    "c5": dict(model="mm_ConvNeXt", kind="convnext_nano.d1h_in1k", batch=1024, cpu_sample=128, cpu_batch=64,
               label="C5 multimodal ConvNeXt-nano training step (forward + BCE + backward + AdamW, NCCL gradient "
                     "all-reduce overlapped with backward), 63x63x3 triplet + 25 metadata per alert"),
}


#: kernel families that are dense contractions on the tcgen05 tensor cores (K4/K5 GEMMs incl. the fused
#: fc1-GELU-fc2 kernel, the tensor-core stem, MaxViT 1x1 / Linear) -> tensor roofline; everything else -> HBM roofline
TENSOR_FAMILIES = ("gemm", "mlp_fused", "mv_expand", "mv_project", "mv_qkv", "mv_proj", "mv_fc", "mv_stem2",
                   "mv_shortcut", "t_wgrad_tc")


def kernel_bound(name: str, precision: str) -> str:
    if name == "t_gemm_tn":        # fp32 CUDA-core split-K GEMM of the training path (tiny head / metadata Linears)
        return "hbm"
    return "tensor" if precision == "bf16" and any(t in name for t in TENSOR_FAMILIES) else "hbm"


def add_roofline_fractions(kernels: dict, pk: dict, precision: str) -> None:
    """Per kernel family: which roofline SURVEY.md 8d files it under and the fraction of the measured peak it reaches
    (algorithmic bytes or flops per launch / CUDA-event time per launch)."""
    for name, k in kernels.items():
        k["bound"] = kernel_bound(name, precision)
        k["frac"] = k["tflops"] / pk["tf_sust"] if k["bound"] == "tensor" else k["gbs"] / pk["hbm"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="alerts per GPU per step (default: the workload's)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-sample", type=int, default=None, help="alerts in the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="c5: issue the training step eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    args.batch = args.batch or wl["batch"]
    args.cpu_sample = args.cpu_sample or wl["cpu_sample"]
    return args


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]),
                    tf_sust=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback")


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML is polled from a
    thread every 5 ms (a default run's timed region is ~0.1 s, too short for `nvidia-smi -lms`); falls back to one
    nvidia-smi query when pynvml is unavailable."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.sm, self.mask, self.stop_flag, self.t, self.h, self.nv = index, [], 0, False, None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            # torch device index -> NVML handle (honours CUDA_VISIBLE_DEVICES through the PCI bus id)
            bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(torch.cuda.get_device_properties(index), "pci_bus_id") else None
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if int(pynvml.nvmlDeviceGetPciInfo(h).bus) == int(bus):
                        self.h = h
            self.nv = pynvml
            self.max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv is None:
            return
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def stop(self):
        if self.nv is None:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=10).stdout.split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "samples": 1, "reasons": [],
                        "note": "pynvml unavailable: one nvidia-smi query after the timed region"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["clock query unavailable"]}
        self.stop_flag = True
        self.t.join(timeout=2)
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max,
                "samples": len(self.sm), "reasons": reasons}


# ----------------------------------------------------------------------------------------------------------
def cpu_port(sample_alerts: int, cfg, sd_np, threads: int):
    """inference_example.py:62-91 on the CPU oracle: float32, eval, no_grad, DataLoader(batch 64, shuffle=False).
    Returns alerts/s over `sample_alerts` synthetic alerts (after one warm-up batch)."""
    from btsbot_b200 import synth, utils
    if "MaxViT" in cfg["model_name"]:
        from oracle import maxvit_oracle as O
    else:
        from oracle import convnext_oracle as O
    from torch.utils.data import DataLoader
    torch.set_num_threads(threads)
    sd = synth.to_torch(sd_np)
    trip = synth.make_triplets(sample_alerts, start=0)
    meta = synth.make_metadata(sample_alerts, start=0)
    img = torch.from_numpy(np.ascontiguousarray(np.transpose(trip.astype(np.float32), (0, 3, 1, 2))))
    ds = utils.FlexibleDataset(images=img, metadata=torch.from_numpy(meta), labels=torch.zeros(sample_alerts, dtype=torch.long))
    dl = DataLoader(ds, batch_size=64, shuffle=False, num_workers=0)
    nw = 8 if "MaxViT" in cfg["model_name"] else 64
    O.forward(sd, cfg, img[:nw], torch.from_numpy(meta[:nw]))       # warm-up
    t0 = time.perf_counter()
    n = 0
    for ib, mb, _ in dl:
        logits = O.forward(sd, cfg, ib, mb)
        _ = torch.sigmoid(logits).round()
        n += ib.shape[0]
    dt = time.perf_counter() - t0
    return n / dt, dt


def run_reference(args, rank):
    if rank != 0:
        return
    from btsbot_b200 import synth
    wl = WORKLOADS[args.workload]
    cfg = synth.canonical_config(wl["model"], wl["kind"])
    sd = synth.make_state_dict(cfg, seed=2)
    threads = os.cpu_count() or 1
    per_step = {"c3": 1024, "c4": 64, "c5": 128}[args.workload]      # bounded sample of the workload per step
    port = cpu_train_port if args.workload == "c5" else cpu_port
    for _ in range(max(1, min(args.warmup, 2))):
        port(per_step // 4, cfg, sd, threads)
    t_total, n_total = 0.0, 0
    for _ in range(args.steps):
        rate, dt = port(per_step, cfg, sd, threads)
        t_total += dt
        n_total += per_step
    value = n_total / t_total
    line = {
        "impl": "reference", "metric": "alerts/sec", "value": value, "unit": "alerts/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["label"] + " (CPU port of the reference path)",
                   "model_kind": wl["kind"], "alerts_per_step": per_step, "batch": 64},
        "cpu_baseline": {"value": value, "unit": "alerts/s", "cores": threads, "kind": "port",
                         "sample": f"{per_step} synthetic alerts per step in batches of 64, torch {torch.__version__} CPU fp32, "
                                   f"{threads} threads; reference itself not runnable offline (timm missing)"},
        "e2e": {"value": value, "unit": "alerts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
def cpu_train_port(sample_alerts: int, cfg, sd_np, threads: int, batch: int = 64):
    """train.py:496-547 on the CPU oracle: torch autograd through the functional fp32 forward + torch.optim.AdamW."""
    from btsbot_b200 import synth
    from oracle import convnext_oracle as O
    torch.set_num_threads(threads)
    sd = {k: v.clone() for k, v in synth.to_torch(sd_np).items()}
    params = []
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
            params.append(v)
    opt = torch.optim.AdamW(params, lr=1e-4, betas=(0.9, 0.999))
    n = max(batch, sample_alerts // batch * batch)
    img = torch.from_numpy(np.ascontiguousarray(synth.make_triplets(n, start=0).transpose(0, 3, 1, 2)))
    meta = torch.from_numpy(synth.make_metadata(n, start=0))
    lab = torch.from_numpy(synth.make_labels(n, start=0)).float().unsqueeze(1)
    cfg = dict(cfg, meta_dropout=0.0, comb_dropout=0.0)         # the oracle's dropout layers are identity

    def one(lo):
        opt.zero_grad()
        lg = O.forward_train(sd, cfg, img[lo:lo + batch], meta[lo:lo + batch])
        loss = torch.nn.functional.binary_cross_entropy_with_logits(lg, lab[lo:lo + batch])
        loss.backward()
        opt.step()
    one(0)                                                      # warm-up
    t0 = time.perf_counter()
    for lo in range(0, n, batch):
        one(lo)
    dt = time.perf_counter() - t0
    return n / dt, dt


def run_c5(args, rank, local_rank, world, dev, cfg, sd_np, wl):
    """BASELINE config C5: one training step per `step`, data parallel over the ranks."""
    import torch.distributed as dist
    import btsbot_b200 as btsbot
    from btsbot_b200 import synth, _lib
    from btsbot_b200._autograd import BCEWithLogitsLoss, FusedAdamW, GraphedTrainStep
    from btsbot_b200.parallel import DistributedDataParallel
    B = args.batch
    model = getattr(btsbot, wl["model"])(cfg)
    model.load_state_dict(synth.to_torch(sd_np), strict=True)
    model = model.to(dev).train()
    ddp = DistributedDataParallel(model, bucket_mb=8.0)
    use_graph = not args.no_graph and (world == 1 or os.environ.get("BTSB_GRAPH_DDP", "1") != "0")
    opt = FusedAdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.999), capturable=use_graph)
    loss_fn = BCEWithLogitsLoss(pos_weight=torch.tensor([1.0]))
    pool, nres = 1024, 2
    trip = synth.make_triplets(pool, start=(rank * B) % (1 << 20))
    meta_pool = synth.make_metadata(pool, start=(rank * B) % (1 << 20))
    lab_pool = synth.make_labels(pool, start=(rank * B) % (1 << 20)).astype(np.float32)
    img_pool = torch.from_numpy(np.ascontiguousarray(trip.astype(np.float32).transpose(0, 3, 1, 2)))
    g = torch.Generator(device="cpu").manual_seed(99 + rank)
    host, res = [], []
    for r in range(nres):
        idx = torch.randint(0, pool, (B,), generator=g)
        h = (img_pool[idx].contiguous().pin_memory(), torch.from_numpy(meta_pool)[idx].contiguous().pin_memory(),
             torch.from_numpy(lab_pool)[idx].unsqueeze(1).contiguous().pin_memory())
        host.append(h)
        res.append(tuple(t.to(dev) for t in h))
    torch.cuda.synchronize()

    def eager_step(img, meta, lab):
        ddp.zero_grad()
        loss = loss_fn(ddp(image_input=img, metadata_input=meta), lab)
        loss.backward()
        opt.step()
        return loss

    # the whole step (~360 kernels, a third of them a few microseconds long; with N > 1 also the NCCL bucket all-reduces
    # on the side stream: 212 k vs 158 k alerts/s at N = 2) is captured once in a CUDA graph and replayed
    stepper = GraphedTrainStep(ddp, opt, loss_fn, example=res[0], warmup=2) if use_graph else None
    train_step = stepper if use_graph else eager_step

    def step(i):
        return train_step(*res[i % nres])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    launches = _lib.launch_count() - n0
    if use_graph:
        launches = stepper.kernels_per_step * args.steps      # kernels inside the replayed graph
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    # e2e: the step a train.py user runs -- pinned host batch -> device (inside the timed region), step, loss -> host
    loss_host = torch.empty((1,), dtype=torch.float32).pin_memory()

    def e2e_step(i):
        h = host[i % nres]
        if use_graph:
            loss = stepper(*h)                                   # pinned host -> static device buffers, then one replay
        else:
            loss = eager_step(*(t.to(dev, non_blocking=True) for t in h))
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
    for i in range(2):
        e2e_step(i)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        e2e_step(i)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    prof = _lib.KernelProfiler()
    _lib.profiler = prof
    for i in range(min(args.steps, 5)):
        (stepper._step() if use_graph else step(i))             # per-kernel events need eager launches
    _lib.profiler = None
    kern = prof.summary()
    nprof = min(args.steps, 5)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    if stepper is not None:
        stepper.release()                       # a graph that captured NCCL work must go before its communicator
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    total = world * B * args.steps
    tot_ms = sum(v["ms"] for v in kern.values())
    kernels = {}
    for name, a in sorted(kern.items(), key=lambda kv: -kv[1]["ms"]):
        per = a["ms"] / a["launches"]
        kernels[name] = {"launches_per_step": a["launches"] / nprof, "ms_per_launch": per,
                         "gbs": a["bytes"] / a["launches"] / (per * 1e-3) / 1e9,
                         "tflops": a["flops"] / a["launches"] / (per * 1e-3) / 1e12, "share": a["ms"] / tot_ms}
    # the roofline entry is the tensor-core GEMM family with the largest share (the three GEMMs of every Linear are
    # 92 % of the step's FLOPs); the element-wise fp32 kernels around them are listed in `kernels`
    add_roofline_fractions(kernels, pk, args.precision)
    tc_names = [k for k in kernels if k in ("t_gemm_tc", "t_wgrad_tc")]
    top = tc_names[0] if tc_names else next(iter(kernels))
    tk = kernels[top]
    if tc_names:
        roof = {"kernel": top, "bound": "tensor", "achieved": tk["tflops"], "peak": pk["tf_sust"], "unit": "TFLOP/s",
                "frac": tk["tflops"] / pk["tf_sust"], "traffic": None,
                "peak_source": pk["source"] + " (sustained bf16 cuBLAS)", "share_of_step": tk["share"],
                "note": "average over every launch of this family in a step (all layer shapes)"}
    else:
        roof = {"kernel": top, "bound": "hbm", "achieved": tk["gbs"], "peak": pk["hbm"], "unit": "GB/s",
                "frac": tk["gbs"] / pk["hbm"], "traffic": None, "peak_source": pk["source"]}
    roof["kernel_time_sum_ms_per_step"] = tot_ms / nprof
    cpu = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, dt = cpu_train_port(args.cpu_sample, dict(cfg), sd_np, threads)
        cpu = {"value": rate, "unit": "alerts/s", "cores": threads, "kind": "port",
               "sample": f"{args.cpu_sample} synthetic alerts in training steps of 64 ({dt:.1f} s): CPU oracle fp32 + torch "
                         f"autograd + torch.optim.AdamW, torch {torch.__version__}, {threads} threads"}
    in_bytes = B * (ALERT_IN_BYTES + 4)
    line = {
        "metric": "alerts/sec", "value": total / (ms * 1e-3), "unit": "alerts/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": wl["label"], "model_kind": wl["kind"], "alerts_per_gpu_per_step": B,
                   "global_batch": world * B, "optimizer": "AdamW (multi-tensor kernel)",
                   "launch": "one CUDA graph replay per step" if use_graph else "eager (one launch per kernel)",
                   "parallelism": f"dp{world}: NCCL all-reduce (avg) of 8 MB gradient buckets on a side stream",
                   "l2_policy": f"a step touches > 3 GB of activations (L2 126 MB); {nres} resident batches rotated"},
        "clocks": clocks,
        "e2e": {"value": total / (ms_e2e * 1e-3), "unit": "alerts/s", "h2d_bytes_per_step": world * in_bytes,
                "d2h_bytes_per_step": world * 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "roofline": roof, "kernels": kernels, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    import btsbot_b200 as btsbot
    from btsbot_b200 import synth, _lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    # NUMA: allocate the pinned host batches next to this rank's GPU (restored before the CPU baseline leg)
    from btsbot_b200.parallel import bind_to_device_numa
    prev_affinity = bind_to_device_numa(local_rank) if args.workload != "c5" else None

    wl = WORKLOADS[args.workload]
    cfg = dict(synth.canonical_config(wl["model"], wl["kind"]), precision=args.precision)
    sd_np = synth.make_state_dict(cfg, seed=2)
    if args.workload == "c5":
        return run_c5(args, rank, local_rank, world, dev, cfg, sd_np, wl)
    model = getattr(btsbot, wl["model"])(cfg)
    model.load_state_dict(synth.to_torch(sd_np), strict=True)
    model = model.to(dev).eval()

    # ---- synthetic inputs: a pool of unique index-keyed alerts for this rank's shard, tiled on device ----------
    pool = 2048
    shard0 = rank * B                                        # this rank's index range starts here
    trip_pool = synth.make_triplets(pool, start=shard0 % (1 << 20))
    meta_pool = synth.make_metadata(pool, start=shard0 % (1 << 20))
    nres = 2                                                 # distinct resident batches rotated across steps
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    res_img, res_meta, host_trip, host_meta = [], [], [], []
    tp = torch.from_numpy(trip_pool).to(dev)
    mp = torch.from_numpy(meta_pool).to(dev)
    for r in range(nres):
        idx = torch.randint(0, pool, (B,), generator=g).to(dev)
        hwc = tp[idx].contiguous()                           # [B,63,63,3] fp32 HWC (what a user holds)
        res_img.append(btsbot.alert_utils.triplets_to_model_input(hwc))          # K1 -> [B,3,63,63] resident
        res_meta.append(mp[idx].contiguous())
        host_trip.append(hwc.cpu().pin_memory())
        host_meta.append(res_meta[-1].cpu().pin_memory())
    del tp, mp
    torch.cuda.synchronize()

    def step(i):
        with torch.no_grad():
            return model(image_input=res_img[i % nres], metadata_input=res_meta[i % nres])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        out = step(i)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- timed region: exactly K steps, device-resident inputs -------------------------------------------------
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        out = step(i)
    e1.record()
    barrier()
    launches = _lib.launch_count() - n0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: public API from pinned host buffers, H2D + K1 + forward + D2H inside the timed region --------------
    out_host = torch.empty((B, 1), dtype=torch.float32).pin_memory()

    # the public bulk-scoring call (btsbot_b200.parallel.AlertScorer): pinned host arrays in, logits out; the H2D copy
    # of micro-batch i+1 runs on a copy stream while the kernels of micro-batch i execute
    from btsbot_b200.parallel import AlertScorer
    scorer = AlertScorer(model, return_scores=False)

    def e2e_step(i):
        lg = scorer(host_trip[i % nres], host_meta[i % nres])
        out_host.copy_(lg.view(-1, 1), non_blocking=True)
    for i in range(2):
        e2e_step(i)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        e2e_step(i)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    # ---- instrumented replay of the same K steps: per-kernel CUDA-event timing ---------------------------------
    prof = _lib.KernelProfiler()
    _lib.profiler = prof
    for i in range(args.steps):
        step(i)
    _lib.profiler = None
    kern = prof.summary()

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    total_alerts = world * B * args.steps
    value = total_alerts / (ms * 1e-3)
    e2e_value = total_alerts / (ms_e2e * 1e-3)
    kernels = {}
    for name, a in sorted(kern.items(), key=lambda kv: -kv[1]["ms"]):
        per = a["ms"] / a["launches"]
        kernels[name] = {"launches_per_step": a["launches"] / args.steps, "ms_per_launch": per,
                         "gbs": a["bytes"] / a["launches"] / (per * 1e-3) / 1e9,
                         "tflops": a["flops"] / a["launches"] / (per * 1e-3) / 1e12,
                         "share": a["ms"] / sum(v["ms"] for v in kern.values())}
    top = next(iter(kernels))
    tk = kernels[top]
    # dense contractions (K4/K5 GEMMs incl. the fused fc1-GELU-fc2 kernel, MaxViT 1x1 / Linear) are
    # reported against the tensor roofline, everything else (dw conv + LN, LN, preprocessing ...) against HBM
    tensor_bound = kernel_bound(top, args.precision) == "tensor"
    add_roofline_fractions(kernels, pk, args.precision)
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures at 8192 alerts
    # per launch (profiles/r01f/*.summary.txt, profiles/r01g/*.summary.txt), scaled to this run's batch
    traffic = NCU_TRAFFIC_8192.get(top)
    if traffic is not None:
        traffic = traffic * B / 8192.0
    if tensor_bound:
        roof = {"kernel": top, "bound": "tensor", "achieved": tk["tflops"], "peak": pk["tf_sust"], "unit": "TFLOP/s",
                "frac": tk["tflops"] / pk["tf_sust"], "traffic": traffic,
                "peak_source": pk["source"] + " (sustained bf16 cuBLAS: kernel timed inside a long step)",
                "hbm_gbs": tk["gbs"], "hbm_frac": tk["gbs"] / pk["hbm"]}
    else:
        roof = {"kernel": top, "bound": "hbm", "achieved": tk["gbs"], "peak": pk["hbm"], "unit": "GB/s",
                "frac": tk["gbs"] / pk["hbm"], "traffic": traffic, "peak_source": pk["source"] + " (copy bandwidth)"}
    roof["algorithmic_bytes"] = tk["gbs"] * 1e9 * tk["ms_per_launch"] * 1e-3
    roof["kernel_time_sum_ms_per_step"] = sum(v["ms"] for v in kern.values()) / args.steps

    cpu = None
    if prev_affinity is not None:
        os.sched_setaffinity(0, prev_affinity)              # the CPU baseline uses every host core
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, dt = cpu_port(args.cpu_sample, dict(cfg), sd_np, threads)
        cpu = {"value": rate, "unit": "alerts/s", "cores": threads, "kind": "port",
               "sample": f"{args.cpu_sample} synthetic alerts in batches of 64 ({dt:.1f} s), CPU oracle fp32, "
                         f"torch {torch.__version__}, {threads} threads"}

    line = {
        "metric": "alerts/sec", "value": value, "unit": "alerts/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": wl["label"],
                   "model_kind": wl["kind"], "alerts_per_gpu_per_step": B, "global_alerts_per_step": world * B,
                   "sharding": "contiguous index ranges, no data-path collective",
                   "l2_policy": f"inputs larger than L2 ({B * ALERT_IN_BYTES / 1e6:.0f} MB per step), "
                                f"{nres} resident batches rotated"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "alerts/s", "h2d_bytes_per_step": world * B * ALERT_IN_BYTES,
                "d2h_bytes_per_step": world * B * 4, "ms_per_step": ms_e2e / args.steps,
                "h2d_gbs_per_gpu": B * ALERT_IN_BYTES / (ms_e2e / args.steps * 1e-3) / 1e9,
                "numa_bound": prev_affinity is not None},
        "gpu_launches": int(launches),
        "roofline": roof,
        "kernels": kernels,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
