#!/usr/bin/env python
"""Per-kernel timing at the C3 shapes (8192 alerts): python scripts/kbench.py [--batch 8192] [--reps 10] [--only substr]
Each kernel is launched `reps` times on rotating buffer sets (working set > L2 where the shape allows) between CUDA events."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from btsbot_b200 import ops, _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8192)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--only", default="")
args = ap.parse_args()
B = args.batch
dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu").manual_seed(0)
bf = torch.bfloat16


def rnd(*shape, dtype=bf, scale=1.0):
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(dev)


def timeit(name, fn, nsets, flops=0.0, nbytes=0.0):
    if args.only and args.only not in name:
        return
    for i in range(3):
        fn(i % nsets)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.reps):
        fn(i % nsets)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    print(f"{name:28s} {ms*1e3:9.1f} us   {flops/ms/1e9:8.1f} TF/s  {nbytes/ms/1e6:8.0f} GB/s", flush=True)


def mlp(C, HW):
    M = B * HW
    ys = [rnd(M, C) for _ in range(2)]; rs = [rnd(M, C) for _ in range(2)]
    w1, w2 = rnd(4 * C, C, scale=C ** -0.5), rnd(C, 4 * C, scale=(4 * C) ** -0.5)
    b1, b2, gm = rnd(4 * C, dtype=torch.float32, scale=0.1), rnd(C, dtype=torch.float32, scale=0.1), rnd(C, dtype=torch.float32)
    timeit(f"mlp_fused C={C} HW={HW}", lambda i: ops.mlp_fused(ys[i], rs[i], w1, b1, w2, b2, gm), 2,
           flops=16.0 * M * C * C, nbytes=2.0 * 3 * M * C)


def fc(C, HW):
    M = B * HW
    ys = [rnd(M, C) for _ in range(2)]; rs = [rnd(M, C) for _ in range(2)]
    hid = [rnd(M, 4 * C) for _ in range(2)]
    w1, w2 = rnd(4 * C, C, scale=C ** -0.5), rnd(C, 4 * C, scale=(4 * C) ** -0.5)
    b1, b2, gm = rnd(4 * C, dtype=torch.float32, scale=0.1), rnd(C, dtype=torch.float32, scale=0.1), rnd(C, dtype=torch.float32)
    timeit(f"gemm fc1+gelu C={C} HW={HW}", lambda i: ops.gemm(ys[i], w1, b1, L.EPI_BIAS_GELU), 2,
           flops=8.0 * M * C * C, nbytes=2.0 * (M * C + 4 * M * C + 4 * C * C))
    timeit(f"gemm fc1+bias C={C} HW={HW}", lambda i: ops.gemm(ys[i], w1, b1, L.EPI_BIAS), 2,
           flops=8.0 * M * C * C, nbytes=2.0 * (M * C + 4 * M * C + 4 * C * C))
    timeit(f"gemm fc2+res C={C} HW={HW}", lambda i: ops.gemm(hid[i], w2, b2, L.EPI_SCALE_RES, gm, rs[i]), 2,
           flops=8.0 * M * C * C, nbytes=2.0 * (4 * M * C + 2 * M * C + 4 * C * C))


def dw(C, S):
    M = B * S * S
    xs = [rnd(M, C) for _ in range(2)]
    w49, bias = rnd(49, C, dtype=torch.float32, scale=0.14), rnd(C, dtype=torch.float32, scale=0.1)
    lw, lb = rnd(C, dtype=torch.float32), rnd(C, dtype=torch.float32)
    timeit(f"dwln {S}x{S}x{C}", lambda i: ops.dwln(xs[i], B, S, S, w49, bias, lw, lb), 2,
           flops=2.0 * 49 * M * C, nbytes=4.0 * M * C)


def lnp(C, S):
    xs = [rnd(B * S * S, C) for _ in range(2)]
    lw, lb = rnd(C, dtype=torch.float32), rnd(C, dtype=torch.float32)
    so = (S - 2) // 2 + 1
    timeit(f"lnpatch {S}x{S}x{C}", lambda i: ops.lnpatch(xs[i], B, S, S, lw, lb), 2,
           nbytes=2.0 * B * C * (S * S + 4 * so * so))


def head():
    from btsbot_b200 import synth, _engine as E
    cfg = synth.canonical_config("mm_ConvNeXt", "convnext_nano.d1h_in1k")
    sd = {k: v.to(dev) for k, v in synth.to_torch(synth.make_state_dict(cfg, seed=2)).items()}
    hw = E.HeadWeights(sd, meta_prefix="metadata_branch.", head_prefix="combined_head.")
    feat = rnd(B, 640); meta = rnd(B, 25, dtype=torch.float32)
    timeit("meta_head", lambda i: E.head_forward(hw, feat, meta, B), 1, flops=2.0 * 119e3 * B, nbytes=B * (1280 + 104.0))


print(f"kbench: batch {B}, reps {args.reps}, BTSB_MLP_V1={os.environ.get('BTSB_MLP_V1', '')}")
mlp(80, 225); mlp(160, 49); mlp(64, 225); mlp(128, 49); mlp(320, 9); mlp(256, 9)
fc(320, 9); fc(640, 1)
dw(80, 15); dw(160, 7); dw(320, 3)
lnp(80, 15); lnp(160, 7)
head()
