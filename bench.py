#!/usr/bin/env python
"""bench.py -- alerts/sec of the multimodal ConvNeXt scoring hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference|reference-gpu] [--batch B]
                    [--precision bf16|fp32] [--workload c3|c2|c4|c5] [--no-extras] [--alerts N]

Headline workload (BASELINE.json configs[2], "C3"): bulk scoring of synthetic alerts with mm_ConvNeXt / convnext_nano
(random-init weights, 25 metadata columns), index-range sharded over ranks with NO collective on the data path.
A "step" is one forward pass of every rank over one micro-batch of B alerts ([B,3,63,63] fp32 + [B,25] fp32,
390 MB at B=8192 -> larger than the 126 MB L2, and consecutive steps rotate over distinct resident batches).

  value    : alerts/s, inputs already resident in HBM, CUDA events around exactly K steps, max over ranks
  e2e      : same metric through the public API (parallel.AlertScorer) from pinned HOST buffers: H2D copy of the HWC
             triplets + metadata (copy stream, overlapping the previous micro-batch's kernels), K1 layout kernel, model
             forward, D2H of the logits -- all inside the timed region; `pcie` = the bare H2D rate of the same buffers
             measured right after, every rank copying at once (the ceiling e2e can reach for this input format)
  roofline : dominant kernel (largest share of the step) timed per launch with CUDA events in an instrumented
             replay of the same K steps right after the timed region (events perturb launches, so `value`
             comes from the un-instrumented pass; `kernels` lists every kernel family incl. K1 for cross-checking);
             `frac` uses the sustained cuBLAS peak, `frac_burst` the burst one; `traffic` is parsed from the newest
             committed `ncu --set full` summary under profiles/ (file named in `traffic_source`)
  sustained: BASELINE's C3 literally -- 1 M alerts (or --alerts) scored back to back, with the clocks seen meanwhile
  cpu_baseline : the CPU oracle (port of the reference path: btsbot/architectures.py glue + restated timm trunk)
             following inference_example.py:62-91 (DataLoader batch 64, num_workers 4, fp32, eval/no_grad) on a bounded
             sample, plus the 39 shipped example alerts (BASELINE config C1)

Sub-records of the default line (same N, a few seconds each; `--no-extras` skips them):
  c5   : BASELINE configs[4], one TRAINING step (zero_grad, forward, BCE, backward, AdamW; train.py:496-547) of
         mm_ConvNeXt-nano on 1024 alerts per GPU in mixed precision, CUDA-graph replay, NCCL bucket all-reduce on a side
         stream inside the graph: value, ms_per_step, e2e; for N > 1 also allreduce_ms (the buckets alone), compute_ms (the
         same step of an un-wrapped replica, no collectives), exposed_ms = ms_per_step - compute_ms and
         overlap_frac = 1 - exposed_ms / allreduce_ms
  c4   : BASELINE configs[3], multimodal MaxViT-tiny-rw-224, batch 4096 per GPU, bf16
  c2   : BASELINE configs[1], image-only ConvNeXt-nano, batch 1024 per GPU
  fp32 : C3 in the package's default precision (the 1e-4 mode)
  gpu_library_baseline (N = 1 only): the same model as eager PyTorch on this GPU (cuDNN / cuBLAS / ATen: what the
         reference runs when given a GPU) -- fp32 with TF32 off, TF32 on, bf16 autocast + channels_last -- and, per
         kernel family of `kernels`, the time of the equivalent library ops (`vs_library` = library time / ours)

`--workload c2|c4|c5` makes that configuration the headline line instead.
`--impl reference` times the CPU port alone with all host threads (the reference itself cannot run offline: timm is
not installable); `--impl reference-gpu` prints the gpu_library_baseline as a line of its own.
"""
import argparse
import glob
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALERT_IN_BYTES = 63 * 63 * 3 * 4 + 25 * 4

WORKLOADS = {
    "c3": dict(model="mm_ConvNeXt", kind="convnext_nano.d1h_in1k", batch=8192, cpu_sample=8192,
               label="C3 multimodal ConvNeXt-nano bulk scoring, 63x63x3 triplet + 25 metadata per alert"),
    "c2": dict(model="ConvNeXt", kind="convnext_nano.d1h_in1k", batch=1024, cpu_sample=1024, graph=True,
               label="C2 image-only ConvNeXt-nano scoring, 63x63x3 triplet per alert, batch 1024 "
                     "(forward replayed as one CUDA graph: config infer_cuda_graph)"),
    "c4": dict(model="mm_MaxViT", kind="maxvit_tiny_rw_224.sw_in1k", batch=4096, cpu_sample=64,
               label="C4 multimodal MaxViT-tiny-rw-224 scoring (bilinear 63->224, MBConv, window+grid attention), "
                     "63x63x3 triplet + 25 metadata per alert"),
    "c5": dict(model="mm_ConvNeXt", kind="convnext_nano.d1h_in1k", batch=1024, cpu_sample=128,
               label="C5 multimodal ConvNeXt-nano training step (forward + BCE + backward + AdamW, NCCL gradient "
                     "all-reduce overlapped with backward), 63x63x3 triplet + 25 metadata per alert"),
}

#: kernel families that are dense contractions on the tcgen05 tensor cores (K4/K5 GEMMs incl. the fused
#: fc1-GELU-fc2 kernel, the tensor-core stem, MaxViT 1x1 / Linear / attention) -> tensor roofline; the rest -> HBM
TENSOR_FAMILIES = ("gemm", "mlp_fused", "mv_expand", "mv_project", "mv_qkv", "mv_proj", "mv_fc", "mv_stem2",
                   "mv_shortcut", "mv_attn_tc", "t_wgrad_tc")

#: kernel family -> stem of its `ncu --set full` summary under profiles/<visit>/ (scripts/ncu_summary.py output)
#: fraction of the 49 taps that fall inside the map: what the conv kernels actually multiply
FP32_VALID_TAPS = {"dwln_15": 38.4 / 49.0, "dwln_7": 27.9 / 49.0}
NCU_SUMMARY = {"mlp_fused_320": "mlp2_320", "mlp_fused_80": "mlp2_80", "mlp_fused_160": "mlp2_160",
               "dwln_15x80": "dwln15", "gemm_fc1_640": "fc1_640", "gemm_fc2_640": "fc2_640"}


def kernel_bound(name: str, precision: str) -> str:
    if name == "t_gemm_tn":        # fp32 CUDA-core split-K GEMM of the training path (tiny head / metadata Linears)
        return "hbm"
    return "tensor" if any(t in name for t in TENSOR_FAMILIES) and (precision == "bf16" or "tf32" in name) else "hbm"


def add_roofline_fractions(kernels: dict, pk: dict, precision: str) -> None:
    """Per kernel family: which roofline SURVEY.md 8d files it under and the fraction of the measured peak it reaches
    (algorithmic bytes or flops per launch / CUDA-event time per launch)."""
    for name, k in kernels.items():
        k["bound"] = kernel_bound(name, precision)
        k["frac"] = k["tflops"] / pk["tf_sust"] if k["bound"] == "tensor" else k["gbs"] / pk["hbm"]
        # dw7x7 + LN on the 15^2 / 7^2 maps: SURVEY.md files it under HBM (kept in `bound` / `frac`), but on B200 the kernel
        # sits on the FP32 FMA pipe (ncu: top stall math-pipe throttle): also report the valid-tap FMA rate against the
        # nominal pipe peak (SMs x 128 FMA/clk x SM clock; the measured FFMA2 ceiling of this access pattern is 0.77 of it)
        ratio = FP32_VALID_TAPS.get(name.split("x")[0])
        if ratio is not None and k.get("tflops"):
            peak = pk.get("sms", 148) * 128 * 2 * pk.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
            k["fp32_pipe"] = {"tflops_valid_taps": k["tflops"] * ratio, "peak": peak, "frac": k["tflops"] * ratio / peak}


def ncu_traffic(kernel: str, batch: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the newest committed ncu summary
    (captured at 8192 alerts per launch; scaled to this run's batch).  Returns (bytes or None, file or None)."""
    stem = NCU_SUMMARY.get(kernel)
    if stem is None:
        return None, None
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*", stem + ".summary.txt")))
    if not files:
        return None, None
    txt = open(files[-1]).read()
    tot = 0.0
    for key in ("dram_rd", "dram_wr"):
        m = re.search(key + r"=([0-9.eE+-]+)([KMG]?)byte", txt)
        if not m:
            return None, None
        tot += float(m.group(1)) * {"": 1.0, "K": 1e3, "M": 1e6, "G": 1e9}[m.group(2)]
    return tot * batch / 8192.0, os.path.relpath(files[-1], ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="alerts per GPU per step (default: the workload's)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-sample", type=int, default=None, help="alerts in the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the c5 / c4 / c2 / fp32 / library sub-records")
    ap.add_argument("--alerts", type=int, default=1000000, help="size of the sustained bulk-scoring run (BASELINE C3: 1 M)")
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
            prop = torch.cuda.get_device_properties(index)
            bus = prop.pci_bus_id if hasattr(prop, "pci_bus_id") else None
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
# CPU legs (the only places that execute oracle/)
# ----------------------------------------------------------------------------------------------------------
def _oracle_for(cfg):
    if "MaxViT" in cfg["model_name"]:
        from oracle import maxvit_oracle as O
    else:
        from oracle import convnext_oracle as O
    return O


def _oracle_call(O, sd, cfg, img, meta):
    if cfg["model_name"] in ("ConvNeXt", "MaxViT"):
        return O.forward(sd, cfg, img, None)
    return O.forward(sd, cfg, img, meta)


def cpu_port(sample_alerts: int, cfg, sd_np, threads: int, num_workers: int = 4):
    """inference_example.py:62-91 on the CPU oracle: float32, eval, no_grad, DataLoader(batch 64, shuffle=False,
    num_workers=4 as the reference).  Returns alerts/s over `sample_alerts` synthetic alerts (after one warm-up batch)."""
    from btsbot_b200 import synth, utils
    from torch.utils.data import DataLoader
    O = _oracle_for(cfg)
    torch.set_num_threads(threads)
    sd = synth.to_torch(sd_np)
    trip = synth.make_triplets(sample_alerts, start=0)
    meta = synth.make_metadata(sample_alerts, start=0)
    img = torch.from_numpy(np.ascontiguousarray(np.transpose(trip.astype(np.float32), (0, 3, 1, 2))))
    ds = utils.FlexibleDataset(images=img, metadata=torch.from_numpy(meta), labels=torch.zeros(sample_alerts, dtype=torch.long))
    try:
        dl = DataLoader(ds, batch_size=64, shuffle=False, num_workers=num_workers)
        it = iter(dl)
    except Exception:                   # no fork / shared memory in this sandbox: in-process loading
        num_workers = 0
        dl = DataLoader(ds, batch_size=64, shuffle=False, num_workers=0)
        it = iter(dl)
    nw = 8 if "MaxViT" in cfg["model_name"] else 64
    with torch.no_grad():
        _oracle_call(O, sd, cfg, img[:nw], torch.from_numpy(meta[:nw]))       # warm-up
        t0 = time.perf_counter()
        n = 0
        for ib, mb, _ in it:
            logits = _oracle_call(O, sd, cfg, ib, mb)
            _ = torch.sigmoid(logits).round()
            n += ib.shape[0]
        dt = time.perf_counter() - t0
    return n / dt, dt, num_workers


def cpu_example_alerts(threads: int):
    """BASELINE config C1: the 39 shipped example alerts (tests/golden/example_inputs.npz = btsbot/example_data) through
    mm_ConvNeXt / convnext_nano on the CPU oracle, one batch (39 < 64); median of 5 passes after one warm-up."""
    from btsbot_b200 import synth
    from oracle import convnext_oracle as O
    path = os.path.join(ROOT, "tests", "golden", "example_inputs.npz")
    if not os.path.isfile(path):
        return None
    ex = np.load(path)
    cfg = synth.canonical_config("mm_ConvNeXt", "convnext_nano.d1h_in1k")
    sd = synth.to_torch(synth.make_state_dict(cfg, seed=2))
    img = torch.from_numpy(np.ascontiguousarray(ex["triplets"].astype(np.float32).transpose(0, 3, 1, 2)))
    meta = torch.from_numpy(ex["metadata"].astype(np.float32))
    torch.set_num_threads(threads)
    ts = []
    with torch.no_grad():
        O.forward(sd, cfg, img, meta)
        for _ in range(5):
            t0 = time.perf_counter()
            _ = torch.sigmoid(O.forward(sd, cfg, img, meta)).round()
            ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts))
    return {"alerts": int(img.shape[0]), "ms": 1e3 * dt, "value": img.shape[0] / dt, "unit": "alerts/s"}


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
    return n / dt, dt, 0


def run_reference(args, rank):
    if rank != 0:
        return
    from btsbot_b200 import synth
    wl = WORKLOADS[args.workload]
    cfg = synth.canonical_config(wl["model"], wl["kind"])
    sd = synth.make_state_dict(cfg, seed=2)
    threads = os.cpu_count() or 1
    per_step = {"c3": 1024, "c2": 1024, "c4": 64, "c5": 128}[args.workload]      # bounded sample of the workload per step
    port = cpu_train_port if args.workload == "c5" else cpu_port
    for _ in range(max(1, min(args.warmup, 2))):
        port(per_step // 4, cfg, sd, threads)
    t_total, n_total, workers = 0.0, 0, 0
    for _ in range(args.steps):
        rate, dt, workers = port(per_step, cfg, sd, threads)
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
                         "sample": f"{per_step} synthetic alerts per step in batches of 64 (DataLoader num_workers={workers}), "
                                   f"torch {torch.__version__} CPU fp32, {threads} threads; reference itself not runnable "
                                   f"offline (timm missing)"},
        "e2e": {"value": value, "unit": "alerts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
# helpers shared by the GPU legs
# ----------------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dev = torch.device("cuda", self.local_rank)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.world == 1:
            return [float(v) for v in vals]
        import torch.distributed as dist
        t = torch.tensor(list(vals), device=self.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]


def kernel_table(kern: dict, nsteps: int):
    tot = sum(v["ms"] for v in kern.values()) or 1.0
    kernels = {}
    for name, a in sorted(kern.items(), key=lambda kv: -kv[1]["ms"]):
        per = a["ms"] / a["launches"]
        kernels[name] = {"launches_per_step": a["launches"] / nsteps, "ms_per_launch": per,
                         "gbs": a["bytes"] / a["launches"] / (per * 1e-3) / 1e9,
                         "tflops": a["flops"] / a["launches"] / (per * 1e-3) / 1e12, "share": a["ms"] / tot}
    return kernels, tot / nsteps


def h2d_ceiling(ctx, bufs, reps: int = 6):
    """Bare host->device rate of the e2e step's own pinned buffers, every rank copying at the same time (the host's
    PCIe / memory fabric is shared between the GPUs of a box): GB/s per GPU, time = max over ranks."""
    dst = [torch.empty_like(b, device=ctx.dev) for b in bufs]
    nbytes = sum(b.numel() * b.element_size() for b in bufs)
    for d, b in zip(dst, bufs):
        d.copy_(b, non_blocking=True)
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for d, b in zip(dst, bufs):
            d.copy_(b, non_blocking=True)
    e1.record()
    ctx.barrier()
    (ms,) = ctx.max_over_ranks(e0.elapsed_time(e1))
    return nbytes * reps / (ms * 1e-3) / 1e9


# ----------------------------------------------------------------------------------------------------------
# inference legs: C3 (headline), C2, C4, and C3 in fp32
# ----------------------------------------------------------------------------------------------------------
def bench_infer(ctx, wl_name: str, precision: str, B: int, steps: int, warmup: int, *, profile: bool = True,
                sustained_alerts: int = 0, pcie: bool = False):
    """One inference configuration on this rank's shard; returns the measurements (rank 0 assembles the line)."""
    import btsbot_b200 as btsbot
    from btsbot_b200 import synth, _lib
    from btsbot_b200.parallel import AlertScorer
    wl = WORKLOADS[wl_name]
    multimodal = wl["model"].startswith("mm_")
    cfg = dict(synth.canonical_config(wl["model"], wl["kind"]), precision=precision)
    if (wl.get("graph") or os.environ.get("BTSB_BENCH_C3_GRAPH") == "1") and os.environ.get("BTSB_BENCH_GRAPH", "1") != "0":
        cfg["infer_cuda_graph"] = True                       # small batch: the host paces ~40 launches per forward
    sd_np = synth.make_state_dict(cfg, seed=2)
    model = getattr(btsbot, wl["model"])(cfg)
    model.load_state_dict(synth.to_torch(sd_np), strict=True)
    model = model.to(ctx.dev).eval()
    from btsbot_b200 import _engine

    # ---- synthetic inputs: a pool of unique index-keyed alerts for this rank's shard, tiled on device ----------
    pool = min(2048, max(256, B))
    shard0 = ctx.rank * B                                    # this rank's index range starts here
    trip_pool = synth.make_triplets(pool, start=shard0 % (1 << 20))
    meta_pool = synth.make_metadata(pool, start=shard0 % (1 << 20))
    nres = 2                                                 # distinct resident batches rotated across steps
    g = torch.Generator(device="cpu").manual_seed(1234 + ctx.rank)
    res_img, res_meta, host_trip, host_meta = [], [], [], []
    tp = torch.from_numpy(trip_pool).to(ctx.dev)
    mp = torch.from_numpy(meta_pool).to(ctx.dev)
    for r in range(nres):
        idx = torch.randint(0, pool, (B,), generator=g).to(ctx.dev)
        hwc = tp[idx].contiguous()                           # [B,63,63,3] fp32 HWC (what a user holds)
        res_img.append(btsbot.alert_utils.triplets_to_model_input(hwc))          # K1 -> [B,3,63,63] resident
        res_meta.append(mp[idx].contiguous())
        host_trip.append(hwc.cpu().pin_memory())
        host_meta.append(res_meta[-1].cpu().pin_memory())
    del tp, mp
    torch.cuda.synchronize()

    def step(i):
        with torch.no_grad():
            if multimodal:
                return model(image_input=res_img[i % nres], metadata_input=res_meta[i % nres])
            return model(input_data=res_img[i % nres])

    warm = max(warmup, 3)
    for i in range(warm):
        step(i)
    ctx.barrier()
    sampler = ClockSampler(ctx.local_rank)
    if ctx.rank == 0:
        sampler.start()
    # ---- timed region: exactly K steps, device-resident inputs -------------------------------------------------
    n0 = _lib.launch_count() + _engine.graph_replay_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    ctx.barrier()
    launches = _lib.launch_count() + _engine.graph_replay_launches() - n0      # host-issued + replayed kernels
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if ctx.rank == 0 else None

    # ---- e2e: public API from pinned host buffers, H2D + K1 + forward + D2H inside the timed region --------------
    out_host = torch.empty((B, 1), dtype=torch.float32).pin_memory()
    scorer = AlertScorer(model, return_scores=False)

    def e2e_step(i):
        lg = scorer(host_trip[i % nres], host_meta[i % nres] if multimodal else None)
        out_host.copy_(lg.view(-1, 1), non_blocking=True)
    # warm-up of the e2e pipeline: the scorer calibrates its host-packing fraction on the first call and, if the split looks
    # worthwhile, A/Bs it against the plain copy over its next twelve calls (AlertScorer._probe_record); the decision is
    # taken on the first timed call at the latest (all warm-up work has completed by then)
    for i in range(16 if scorer._pack_ok else 4):
        e2e_step(i)
    ctx.barrier()
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(steps):
        e2e_step(i)
    f1.record()
    ctx.barrier()
    ms_e2e = f0.elapsed_time(f1)
    ms, ms_e2e = ctx.max_over_ranks(ms, ms_e2e)
    # bytes that crossed PCIe per step: the scorer may round the fp32 triplets to bf16 on the host (AlertScorer host_pack)
    # (a fraction of each batch: rows [0, n1) go over as bf16, the rest as fp32)
    pack_f = scorer.last_fraction if scorer._pack_rings else 0.0      # as settled by the calibration + adaptation
    n1 = int(round(pack_f * B))
    in_bytes = 63 * 63 * 3 * (2 * n1 + 4 * (B - n1)) + B * (25 * 4 if multimodal else 0)
    out = {"wl": wl, "cfg": cfg, "sd_np": sd_np, "B": B, "steps": steps, "warmup": warm, "ms": ms, "ms_e2e": ms_e2e,
           "launches": int(launches), "clocks": clocks, "nres": nres, "in_bytes": in_bytes, "multimodal": multimodal,
           "res_bytes": B * (63 * 63 * 3 * 4 + (25 * 4 if multimodal else 0)),      # the device-resident step input
           "host_pack": {"fraction": pack_f, "threads": scorer.pack_threads, "host_bytes_per_step": B * 63 * 63 * 3 * 4,
                         "calibration_ms": None if scorer.last_calibration is None else
                         {"pack": scorer.last_calibration[1], "fp32_copy": scorer.last_calibration[2]},
                         "probe": scorer.last_probe}}
    if pcie:
        out["h2d_gbs_ceiling"] = h2d_ceiling(ctx, [host_trip[0]] + ([host_meta[0]] if multimodal else []))

    # ---- BASELINE C3 literally: `sustained_alerts` alerts back to back (clocks sampled meanwhile) -----------------
    if sustained_alerts:
        nstep = max(1, -(-sustained_alerts // (B * ctx.world)))
        s2 = ClockSampler(ctx.local_rank)
        ctx.barrier()
        if ctx.rank == 0:
            s2.start()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for i in range(nstep):
            step(i)
        g1.record()
        ctx.barrier()
        (ms_s,) = ctx.max_over_ranks(g0.elapsed_time(g1))
        out["sustained"] = {"alerts": nstep * B * ctx.world, "steps": nstep, "value": nstep * B * ctx.world / (ms_s * 1e-3),
                            "unit": "alerts/s", "ms_per_step": ms_s / nstep, "seconds": ms_s * 1e-3,
                            "clocks": s2.stop() if ctx.rank == 0 else None}

    # ---- instrumented replay of the same K steps: per-kernel CUDA-event timing (K1 included) ---------------------
    if profile:
        model._config["infer_cuda_graph"] = False                # per-kernel events need host-issued launches
        prof = _lib.KernelProfiler()
        _lib.profiler = prof
        hwc_dev = host_trip[0].to(ctx.dev)
        for i in range(steps):
            btsbot.alert_utils.triplets_to_model_input(hwc_dev)          # K1 as the e2e path runs it
            step(i)
        _lib.profiler = None
        out["kern"] = prof.summary()
        del hwc_dev
    del model, scorer, res_img, res_meta
    torch.cuda.empty_cache()
    return out


def infer_line(args, ctx, r, pk, extras: dict):
    """Assemble the JSON line of an inference workload from bench_infer's measurements (rank 0)."""
    wl, B, steps = r["wl"], r["B"], r["steps"]
    total_alerts = ctx.world * B * steps
    kernels, ksum = kernel_table(r["kern"], steps)
    add_roofline_fractions(kernels, pk, args.precision)
    # the roofline kernel: largest share among the model's kernels (K1 is listed, but belongs to the e2e path)
    top = next(k for k in kernels if k not in ("crop_norm", "pad_norm"))
    tk = kernels[top]
    traffic, tsrc = ncu_traffic(top, B)
    if tk["bound"] == "tensor":
        roof = {"kernel": top, "bound": "tensor", "achieved": tk["tflops"], "peak": pk["tf_sust"], "unit": "TFLOP/s",
                "frac": tk["tflops"] / pk["tf_sust"], "frac_sustained": tk["tflops"] / pk["tf_sust"],
                "frac_burst": tk["tflops"] / pk["tf_burst"], "peak_burst": pk["tf_burst"], "traffic": traffic,
                "traffic_source": tsrc,
                "peak_source": pk["source"] + " (cuBLAS bf16: sustained for `frac`, burst for `frac_burst`)",
                "hbm_gbs": tk["gbs"], "hbm_frac": tk["gbs"] / pk["hbm"]}
    else:
        roof = {"kernel": top, "bound": "hbm", "achieved": tk["gbs"], "peak": pk["hbm"], "unit": "GB/s",
                "frac": tk["gbs"] / pk["hbm"], "traffic": traffic, "traffic_source": tsrc,
                "peak_source": pk["source"] + " (copy bandwidth)"}
    roof["algorithmic_bytes"] = tk["gbs"] * 1e9 * tk["ms_per_launch"] * 1e-3
    roof["share_of_step"] = tk["share"]
    roof["kernel_time_sum_ms_per_step"] = ksum
    e2e_value = total_alerts / (r["ms_e2e"] * 1e-3)
    h2d = r["in_bytes"] / (r["ms_e2e"] / steps * 1e-3) / 1e9
    e2e = {"value": e2e_value, "unit": "alerts/s", "h2d_bytes_per_step": ctx.world * r["in_bytes"],
           "d2h_bytes_per_step": ctx.world * B * 4, "ms_per_step": r["ms_e2e"] / steps, "h2d_gbs_per_gpu": h2d,
           "numa_bound": extras.pop("numa_bound", False), "host_pack": r.get("host_pack")}
    if "h2d_gbs_ceiling" in r:
        e2e["pcie"] = {"h2d_gbs_per_gpu_bare": r["h2d_gbs_ceiling"], "e2e_over_pcie_peak": h2d / r["h2d_gbs_ceiling"],
                       "how": "the same pinned buffers copied back to back, all ranks at once, max over ranks"}
    line = {
        "metric": "alerts/sec", "value": total_alerts / (r["ms"] * 1e-3), "unit": "alerts/s", "n_gpus": ctx.world,
        "steps": steps, "warmup": r["warmup"], "ms_per_step": r["ms"] / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": wl["label"], "model_kind": wl["kind"], "alerts_per_gpu_per_step": B,
                   "global_alerts_per_step": ctx.world * B,
                   "sharding": "contiguous index ranges, no data-path collective",
                   "l2_policy": f"inputs larger than L2 ({r['res_bytes'] / 1e6:.0f} MB per step), "
                                f"{r['nres']} resident batches rotated" if r["res_bytes"] > 126e6 else
                                f"{r['nres']} resident batches rotated; a step's activations exceed L2"},
        "clocks": r["clocks"], "e2e": e2e, "gpu_launches": r["launches"], "roofline": roof, "kernels": kernels,
    }
    if "sustained" in r:
        line["sustained"] = r["sustained"]
    line.update(extras)
    return line


def compact(r, ctx):
    """Sub-record of an inference configuration: value + e2e, same N."""
    tot = ctx.world * r["B"] * r["steps"]
    return {"workload": r["wl"]["label"], "alerts_per_gpu_per_step": r["B"], "value": tot / (r["ms"] * 1e-3),
            "unit": "alerts/s", "ms_per_step": r["ms"] / r["steps"], "steps": r["steps"], "n_gpus": ctx.world,
            "e2e": {"value": tot / (r["ms_e2e"] * 1e-3), "unit": "alerts/s", "h2d_bytes_per_step": ctx.world * r["in_bytes"],
                    "d2h_bytes_per_step": ctx.world * r["B"] * 4, "host_pack": r.get("host_pack")},
            "gpu_launches": r["launches"], "clocks": r["clocks"]}


# ----------------------------------------------------------------------------------------------------------
# training leg: C5
# ----------------------------------------------------------------------------------------------------------
def bench_train(ctx, precision: str, B: int, steps: int, warmup: int, use_graph: bool = True, profile: bool = True):
    """BASELINE config C5: one training step per `step`, data parallel over the ranks (every rank must call this)."""
    import torch.distributed as dist
    import btsbot_b200 as btsbot
    from btsbot_b200 import synth, _lib
    from btsbot_b200._autograd import BCEWithLogitsLoss, FusedAdamW, GraphedTrainStep
    from btsbot_b200.parallel import DistributedDataParallel
    wl = WORKLOADS["c5"]
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    cfg = dict(synth.canonical_config(wl["model"], wl["kind"]), precision=precision)
    sd_np = synth.make_state_dict(cfg, seed=2)
    model = getattr(btsbot, wl["model"])(cfg)
    model.load_state_dict(synth.to_torch(sd_np), strict=True)
    model = model.to(dev).train()
    ddp = DistributedDataParallel(model, bucket_mb=8.0)
    use_graph = use_graph and (world == 1 or os.environ.get("BTSB_GRAPH_DDP", "1") != "0")
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
    # on the side stream) is captured once in a CUDA graph and replayed
    stepper = GraphedTrainStep(ddp, opt, loss_fn, example=res[0], warmup=2) if use_graph else None
    train_step = stepper if use_graph else eager_step

    def step(i):
        return train_step(*res[i % nres])

    def timed(fn, n):
        ctx.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            fn(i)
        b.record()
        ctx.barrier()
        return a.elapsed_time(b)

    warm = max(warmup, 3)
    for i in range(warm):
        step(i)
    ctx.barrier()
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    ms = timed(step, steps)
    launches = _lib.launch_count() - n0
    if use_graph:
        launches = stepper.kernels_per_step * steps             # kernels inside the replayed graph
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
    ms_e2e = timed(e2e_step, steps)

    # ---- the collective alone, and the step without it (N > 1): how much of the all-reduce hides behind the backward.
    # The collective-free step is a SECOND model replica without the DistributedDataParallel wrapper (no gradient sink, no
    # side stream), captured and replayed exactly like the single-GPU step.  (Capturing the wrapped model again with its
    # collectives switched off made the capturing stream wait on the un-captured side stream and hung the 2-GPU run.)
    comm = None
    if world > 1:
        sink = ddp.sink
        chunks = [sink.flat[lo:hi] for lo, hi in sink.bounds]
        keep = sink.flat.clone()

        def ar_only(_):
            for c in chunks:
                dist.all_reduce(c, op=dist.ReduceOp.AVG)
        for i in range(3):
            ar_only(i)
        ms_ar = timed(ar_only, 10) / 10
        sink.flat.copy_(keep)
        ms_solo = None
        if use_graph and os.environ.get("BTSB_BENCH_SOLO", "1") != "0":
            solo_model = getattr(btsbot, wl["model"])(cfg)
            solo_model.load_state_dict(synth.to_torch(sd_np), strict=True)
            solo_model = solo_model.to(dev).train()
            solo_opt = FusedAdamW(solo_model.parameters(), lr=1e-4, betas=(0.9, 0.999), capturable=True)
            solo = GraphedTrainStep(solo_model, solo_opt, loss_fn, example=res[0], warmup=2)
            for i in range(2):
                solo(*res[i % nres])
            ms_solo = timed(lambda i: solo(*res[i % nres]), steps) / steps
            solo.release()
            del solo, solo_opt, solo_model
            (ms_solo,) = ctx.max_over_ranks(ms_solo)
        (ms_ar,) = ctx.max_over_ranks(ms_ar)
        comm = {"allreduce_ms": ms_ar, "compute_ms": ms_solo,
                "how": "allreduce_ms: the step's 8 MB gradient buckets all-reduced back to back with nothing else running; "
                       "compute_ms: the same step of an un-wrapped replica (no collectives), CUDA-graph replay, same rank",
                "payload_bytes": int(sink.total * 4), "buckets": len(sink.bounds), "bucket_mb": 8.0,
                "payload_dtype": "fp32",
                "algo": os.environ.get("NCCL_ALGO", "nccl default (NVLS / ring chosen by NCCL; see NCCL_DEBUG=INFO)")}

    kern = None
    nprof = min(steps, 5)
    if profile:
        prof = _lib.KernelProfiler()
        _lib.profiler = prof
        for i in range(nprof):
            (stepper._step() if use_graph else step(i))             # per-kernel events need eager launches
        _lib.profiler = None
        kern = prof.summary()
    ms, ms_e2e = ctx.max_over_ranks(ms, ms_e2e)
    if comm is not None and comm["compute_ms"] is not None:
        exposed = max(0.0, ms / steps - comm["compute_ms"])          # what the collectives add to the replayed step
        comm["exposed_ms"] = exposed
        comm["overlap_frac"] = float(min(1.0, max(0.0, 1.0 - exposed / comm["allreduce_ms"]))) if comm["allreduce_ms"] > 0 else None
    elif comm is not None:
        comm["exposed_ms"] = comm["overlap_frac"] = None
    if stepper is not None:
        stepper.release()                       # a graph that captured NCCL work must go before its communicator
    out = {"wl": wl, "cfg": cfg, "sd_np": sd_np, "B": B, "steps": steps, "warmup": warm, "ms": ms, "ms_e2e": ms_e2e,
           "launches": int(launches), "clocks": clocks, "kern": kern, "nprof": nprof, "use_graph": use_graph,
           "nres": nres, "comm": comm, "in_bytes": B * (ALERT_IN_BYTES + 4)}
    del model, ddp, opt, stepper
    torch.cuda.empty_cache()
    return out


def train_compact(r, ctx):
    tot = ctx.world * r["B"] * r["steps"]
    d = {"workload": r["wl"]["label"], "alerts_per_gpu_per_step": r["B"], "global_batch": ctx.world * r["B"],
         "value": tot / (r["ms"] * 1e-3), "unit": "alerts/s", "ms_per_step": r["ms"] / r["steps"], "steps": r["steps"],
         "n_gpus": ctx.world, "launch": "one CUDA graph replay per step" if r["use_graph"] else "eager",
         "e2e": {"value": tot / (r["ms_e2e"] * 1e-3), "unit": "alerts/s", "h2d_bytes_per_step": ctx.world * r["in_bytes"],
                 "d2h_bytes_per_step": ctx.world * 4},
         "gpu_launches": r["launches"], "clocks": r["clocks"]}
    if r["comm"] is not None:
        d.update({k: r["comm"][k] for k in ("allreduce_ms", "compute_ms", "exposed_ms", "overlap_frac")})
        d["collective"] = {k: r["comm"][k] for k in ("payload_bytes", "payload_dtype", "buckets", "bucket_mb", "algo")}
    else:
        d.update({"allreduce_ms": 0.0, "compute_ms": r["ms"] / r["steps"], "exposed_ms": 0.0, "overlap_frac": None})
    return d


def train_line(args, ctx, r, pk, cpu):
    wl, B, steps = r["wl"], r["B"], r["steps"]
    kernels, ksum = kernel_table(r["kern"], r["nprof"])
    # the roofline entry is the tensor-core GEMM family with the largest share (the three GEMMs of every Linear are
    # 92 % of the step's FLOPs); the element-wise fp32 kernels around them are listed in `kernels`
    add_roofline_fractions(kernels, pk, args.precision)
    tc_names = [k for k in kernels if k in ("t_gemm_tc", "t_wgrad_tc")]
    top = tc_names[0] if tc_names else next(iter(kernels))
    tk = kernels[top]
    if tc_names:
        roof = {"kernel": top, "bound": "tensor", "achieved": tk["tflops"], "peak": pk["tf_sust"], "unit": "TFLOP/s",
                "frac": tk["tflops"] / pk["tf_sust"], "frac_burst": tk["tflops"] / pk["tf_burst"], "traffic": None,
                "peak_source": pk["source"] + " (sustained bf16 cuBLAS)", "share_of_step": tk["share"],
                "note": "average over every launch of this family in a step (all layer shapes)"}
    else:
        roof = {"kernel": top, "bound": "hbm", "achieved": tk["gbs"], "peak": pk["hbm"], "unit": "GB/s",
                "frac": tk["gbs"] / pk["hbm"], "traffic": None, "peak_source": pk["source"]}
    roof["kernel_time_sum_ms_per_step"] = ksum
    c = train_compact(r, ctx)
    line = {
        "metric": "alerts/sec", "value": c["value"], "unit": "alerts/s", "n_gpus": ctx.world, "steps": steps,
        "warmup": r["warmup"], "ms_per_step": c["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": wl["label"], "model_kind": wl["kind"], "alerts_per_gpu_per_step": B,
                   "global_batch": ctx.world * B, "optimizer": "AdamW (multi-tensor kernel)", "launch": c["launch"],
                   "parallelism": f"dp{ctx.world}: NCCL all-reduce (avg) of 8 MB gradient buckets on a side stream",
                   "l2_policy": f"a step touches > 3 GB of activations (L2 126 MB); {r['nres']} resident batches rotated"},
        "clocks": r["clocks"],
        "e2e": dict(c["e2e"], ms_per_step=r["ms_e2e"] / steps),
        "gpu_launches": r["launches"], "roofline": roof, "kernels": kernels, "cpu_baseline": cpu,
        "collective": r["comm"],
    }
    return line


# ----------------------------------------------------------------------------------------------------------
# the reference stack on the GPU: eager PyTorch (cuDNN / cuBLAS / ATen) -- "library kernels to beat"
# ----------------------------------------------------------------------------------------------------------
def _time_cuda(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def gpu_library_baseline(dev, wl_name: str, B: int, steps: int = 5):
    """The oracle model (the reference's architectures.py glue over a restated timm trunk, all torch.nn.functional ops)
    moved to the GPU and run eagerly -- the stack the reference itself runs on a GPU (architectures.py:166-171):
    (a) fp32 with TF32 off, (b) TF32 on, (c) bf16 autocast with channels_last input.  alerts/s at B alerts per step."""
    from btsbot_b200 import synth
    wl = WORKLOADS[wl_name]
    cfg = synth.canonical_config(wl["model"], wl["kind"])
    O = _oracle_for(cfg)
    sd = {k: v.to(dev) for k, v in synth.to_torch(synth.make_state_dict(cfg, seed=2)).items()}
    img = torch.from_numpy(np.ascontiguousarray(synth.make_triplets(min(B, 1024), start=0).transpose(0, 3, 1, 2))).to(dev)
    meta = torch.from_numpy(synth.make_metadata(min(B, 1024), start=0)).to(dev)
    rep = -(-B // img.shape[0])
    img, meta = img.repeat(rep, 1, 1, 1)[:B].contiguous(), meta.repeat(rep, 1)[:B].contiguous()
    out = {"alerts_per_step": B, "unit": "alerts/s", "torch": torch.__version__,
           "cudnn": torch.backends.cudnn.version(), "what": "oracle model .cuda(), eager, no_grad"}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        with torch.no_grad():
            for name, tf32 in (("fp32_tf32_off", False), ("fp32_tf32_on", True)):
                torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.backends.cudnn.allow_tf32 = tf32
                ms = _time_cuda(lambda: _oracle_call(O, sd, cfg, img, meta), steps)
                out[name] = {"value": B / (ms * 1e-3), "ms_per_step": ms}
            img_cl = img.contiguous(memory_format=torch.channels_last)
            sd_cl = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd.items()}
            with torch.autocast("cuda", dtype=torch.bfloat16):
                ms = _time_cuda(lambda: _oracle_call(O, sd_cl, cfg, img_cl, meta), steps)
            out["bf16_autocast_channels_last"] = {"value": B / (ms * 1e-3), "ms_per_step": ms}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    return out


def library_kernel_table(dev, B: int, dims=(80, 160, 320, 640), maps=(15, 7, 3, 1), reps: int = 5):
    """Per kernel family of the C3 `kernels` table: the equivalent eager library ops (cuDNN convolutions, cuBLASLt GEMMs,
    ATen LayerNorm / GELU) in bf16, channels_last, on the same shapes -- ms per call at B alerts."""
    import torch.nn.functional as F
    bf = torch.bfloat16
    t = {}
    with torch.no_grad():
        x0 = torch.randn(B, 3, 63, 63, device=dev)
        w = torch.randn(dims[0], 3, 4, 4, device=dev, dtype=bf).contiguous(memory_format=torch.channels_last)
        lw, lb = torch.ones(dims[0], device=dev, dtype=bf), torch.zeros(dims[0], device=dev, dtype=bf)

        def stem():
            y = F.conv2d(x0.to(bf).contiguous(memory_format=torch.channels_last), w, None, stride=4)
            return F.layer_norm(y.permute(0, 2, 3, 1), (dims[0],), lw, lb, 1e-6)
        t["stem_fused"] = _time_cuda(stem, reps)
        for c, s in zip(dims, maps):
            x = torch.randn(B, c, s, s, device=dev, dtype=bf).contiguous(memory_format=torch.channels_last)
            dw = torch.randn(c, 1, 7, 7, device=dev, dtype=bf).contiguous(memory_format=torch.channels_last)
            db = torch.zeros(c, device=dev, dtype=bf)
            g, b = torch.ones(c, device=dev, dtype=bf), torch.zeros(c, device=dev, dtype=bf)
            w1, b1 = torch.randn(4 * c, c, device=dev, dtype=bf) * c ** -0.5, torch.zeros(4 * c, device=dev, dtype=bf)
            w2, b2 = torch.randn(c, 4 * c, device=dev, dtype=bf) * (4 * c) ** -0.5, torch.zeros(c, device=dev, dtype=bf)
            gam = torch.ones(c, device=dev, dtype=bf)

            def dwln():
                y = F.conv2d(x, dw, db, padding=3, groups=c)
                return F.layer_norm(y.permute(0, 2, 3, 1), (c,), g, b, 1e-6)
            rows = torch.randn(B * s * s, c, device=dev, dtype=bf)
            res = torch.randn(B * s * s, c, device=dev, dtype=bf)

            def fc1():
                return F.gelu(F.linear(rows, w1, b1))
            hid = fc1()

            def fc2():
                return torch.addcmul(res, F.linear(hid, w2, b2), gam)

            def mlp():
                return torch.addcmul(res, F.linear(F.gelu(F.linear(rows, w1, b1)), w2, b2), gam)
            t[f"dwln_{s}x{c}"] = _time_cuda(dwln, reps)
            t[f"mlp_fused_{c}"] = _time_cuda(mlp, reps)
            t[f"gemm_fc1_{c}"] = _time_cuda(fc1, reps)
            t[f"gemm_fc2_{c}"] = _time_cuda(fc2, reps)
            del hid, rows, res, x
        # downsample: LayerNorm2d + conv 2x2 / s2 (lnpatch + gemm_down), averaged over the three stage transitions
        tl, tg = [], []
        for (ci, co), s in zip(zip(dims[:-1], dims[1:]), maps[:-1]):
            x = torch.randn(B, s, s, ci, device=dev, dtype=bf)
            g, b = torch.ones(ci, device=dev, dtype=bf), torch.zeros(ci, device=dev, dtype=bf)
            wd = torch.randn(co, ci, 2, 2, device=dev, dtype=bf).contiguous(memory_format=torch.channels_last)
            bd = torch.zeros(co, device=dev, dtype=bf)
            xn = F.layer_norm(x, (ci,), g, b, 1e-6).permute(0, 3, 1, 2)
            tl.append(_time_cuda(lambda: F.layer_norm(x, (ci,), g, b, 1e-6), reps))
            tg.append(_time_cuda(lambda: F.conv2d(xn, wd, bd, stride=2), reps))
        t["lnpatch"], t["gemm_down"] = sum(tl) / len(tl), sum(tg) / len(tg)
        # metadata branch + fusion head (BatchNorm1d eval, 2 + 3 Linears, GELU), fp32 as the reference runs it
        m = torch.randn(B, 25, device=dev)
        feat = torch.randn(B, dims[-1], device=dev)
        bn = [torch.ones(25, device=dev), torch.zeros(25, device=dev), torch.zeros(25, device=dev), torch.ones(25, device=dev)]
        ws = [torch.randn(128, 25, device=dev), torch.randn(128, 128, device=dev), torch.randn(128, dims[-1] + 128, device=dev),
              torch.randn(8, 128, device=dev), torch.randn(1, 8, device=dev)]

        def head():
            e = F.batch_norm(m, bn[2], bn[3], bn[0], bn[1], False)
            e = F.gelu(F.linear(F.gelu(F.linear(e, ws[0])), ws[1]))
            h = F.gelu(F.linear(torch.cat((feat, e), 1), ws[2]))
            return F.linear(F.gelu(F.linear(h, ws[3])), ws[4])
        t["meta_head"] = _time_cuda(head, reps)
        # K1 as the reference does it: astype(float32) + transpose(0,3,1,2) + contiguous
        hwc = torch.randn(B, 63, 63, 3, device=dev)
        t["crop_norm"] = _time_cuda(lambda: hwc.permute(0, 3, 1, 2).contiguous(), reps)
    return t


def run_reference_gpu(args, ctx):
    if ctx.rank != 0:
        return
    torch.cuda.set_device(ctx.local_rank)
    lib = gpu_library_baseline(ctx.dev, args.workload if args.workload != "c5" else "c3", args.batch)
    best = max(lib[k]["value"] for k in ("fp32_tf32_off", "fp32_tf32_on", "bf16_autocast_channels_last"))
    print(json.dumps({"impl": "reference-gpu", "metric": "alerts/sec", "value": best, "unit": "alerts/s", "n_gpus": 1,
                      "higher_is_better": True, "data": "synthetic", "gpu_library_baseline": lib,
                      "note": "eager PyTorch (cuDNN / cuBLAS / ATen) on this GPU; best of the three precisions as `value`"}),
          flush=True)


# ----------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    ctx = Ctx()
    if args.impl == "reference":
        run_reference(args, ctx.rank)
        return
    if args.impl == "reference-gpu":
        run_reference_gpu(args, ctx)
        return

    import torch.distributed as dist
    torch.cuda.set_device(ctx.local_rank)
    if ctx.world > 1:
        dist.init_process_group("nccl", device_id=ctx.dev)
    pk = peaks()
    threads = os.cpu_count() or 1
    # No try/finally around this: after an exception the process group must NOT be torn down (a CUDA graph that captured
    # NCCL work has to be destroyed before its communicator, and on the failure path nobody does that in order -- the
    # __main__ guard prints the traceback and leaves at once).
    run_b200(args, ctx, pk, threads)
    if ctx.world > 1 and dist.is_initialized():
        dist.destroy_process_group()


def run_b200(args, ctx, pk, threads):
    head = args.workload
    extras_on = not args.no_extras
    if head == "c5":
        r = bench_train(ctx, args.precision, args.batch, args.steps, args.warmup, use_graph=not args.no_graph)
        if ctx.rank == 0:
            cpu = None
            if not args.no_cpu_baseline:
                rate, dt, _ = cpu_train_port(args.cpu_sample, dict(r["cfg"]), r["sd_np"], threads)
                cpu = {"value": rate, "unit": "alerts/s", "cores": threads, "kind": "port",
                       "sample": f"{args.cpu_sample} synthetic alerts in training steps of 64 ({dt:.1f} s): CPU oracle fp32 + "
                                 f"torch autograd + torch.optim.AdamW, torch {torch.__version__}, {threads} threads"}
            print(json.dumps(train_line(args, ctx, r, pk, cpu)), flush=True)
        return

    # NUMA: allocate the pinned host batches next to this rank's GPU (restored before the CPU baseline leg)
    from btsbot_b200.parallel import bind_to_device_numa
    prev_affinity = bind_to_device_numa(ctx.local_rank)
    r = bench_infer(ctx, head, args.precision, args.batch, args.steps, args.warmup, profile=True,
                    sustained_alerts=args.alerts if head == "c3" else 0, pcie=True)
    extras = {"numa_bound": prev_affinity is not None}
    if extras_on and head == "c3":
        # every rank runs the same sequence (C5 has collectives); a few seconds each
        sub_c5 = bench_train(ctx, "bf16", WORKLOADS["c5"]["batch"], 10, 3, profile=False)
        sub_c4 = bench_infer(ctx, "c4", "bf16", WORKLOADS["c4"]["batch"], 3, 3, profile=False)
        sub_c2 = bench_infer(ctx, "c2", args.precision, WORKLOADS["c2"]["batch"], 20, 3, profile=False)
        sub_fp32 = bench_infer(ctx, "c3", "fp32", args.batch, 5, 3, profile=False) if args.precision != "fp32" else None
        extras["c5"] = train_compact(sub_c5, ctx)
        extras["c4"] = compact(sub_c4, ctx)
        extras["c2"] = compact(sub_c2, ctx)
        if sub_fp32 is not None:
            extras["fp32"] = compact(sub_fp32, ctx)
    if ctx.rank != 0:
        return
    if prev_affinity is not None:
        os.sched_setaffinity(0, prev_affinity)              # the CPU legs use every host core
    line = infer_line(args, ctx, r, pk, extras)
    if extras_on and ctx.world == 1 and head in ("c3", "c2"):
        lib = gpu_library_baseline(ctx.dev, head, args.batch)
        line["gpu_library_baseline"] = lib
        if head == "c3":
            tab = library_kernel_table(ctx.dev, args.batch)
            for name, k in line["kernels"].items():
                if name in tab:
                    k["library_ms"] = tab[name]
                    k["vs_library"] = tab[name] / k["ms_per_launch"]
            lib["kernel_table_note"] = ("kernels[*].library_ms = the equivalent eager library ops (cuDNN conv, cuBLASLt "
                                        "GEMM, ATen LayerNorm/GELU; bf16, channels_last) at the same shapes; vs_library = "
                                        "library_ms / ms_per_launch (> 1: this repo's kernel is faster)")
            best = max(lib[k]["value"] for k in ("fp32_tf32_off", "fp32_tf32_on", "bf16_autocast_channels_last"))
            lib["value_over_best_library"] = line["value"] / best
    cpu = None
    if not args.no_cpu_baseline:
        rate, dt, workers = cpu_port(args.cpu_sample, dict(r["cfg"]), r["sd_np"], threads)
        cpu = {"value": rate, "unit": "alerts/s", "cores": threads, "kind": "port",
               "sample": f"{args.cpu_sample} synthetic alerts in batches of 64, DataLoader num_workers={workers} "
                         f"({dt:.1f} s), CPU oracle fp32, torch {torch.__version__}, {threads} threads",
               "example_alerts_c1": cpu_example_alerts(threads) if head == "c3" else None}
    line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    try:
        main()
    except BaseException as exc:                 # noqa: BLE001
        if isinstance(exc, SystemExit) and exc.code in (0, None):
            raise
        # a CUDA graph that captured NCCL work must be destroyed before its communicator; after an exception nobody does
        # that in order, and interpreter teardown then hangs in NCCL -- leave at once instead (torchrun reaps the peers)
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        sys.stdout.flush()
        os._exit(1)
