#!/usr/bin/env python
"""Timeline of the fused-MLP pipeline hand-offs in block 0 (btsb_debug_mlp_trace): python scripts/mlp_trace.py [C] [HW]
Prints, per hidden chunk g, clock offsets (relative to the first G1 issue) of: G1/G2 issue by the MMA warp and, for one
epilogue warp of the chunk's group, loop-top / D1-full seen / TMEM load done / GELU done / H-empty seen / H-full arrive."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from btsbot_b200 import ops, _lib as L

C = int(sys.argv[1]) if len(sys.argv) > 1 else 80
HW = int(sys.argv[2]) if len(sys.argv) > 2 else 225
B = 8192
dev = torch.device("cuda", 0)
M = B * HW
g = torch.Generator().manual_seed(0)
bf = torch.bfloat16
y = torch.randn(M, C, generator=g).to(bf).to(dev); res = torch.randn(M, C, generator=g).to(torch.float16).to(dev)    # the fp16 residual stream
w1 = (torch.randn(4 * C, C, generator=g) * C ** -0.5).to(bf).to(dev); w2 = (torch.randn(C, 4 * C, generator=g) * (4 * C) ** -0.5).to(bf).to(dev)
b1 = torch.randn(4 * C, generator=g).to(dev) * 0.1; b2 = torch.randn(C, generator=g).to(dev) * 0.1; gm = torch.randn(C, generator=g).to(dev)
INPLACE = os.environ.get("BTSB_MLP_INPLACE", "0") != "0"      # out == res: the wide kernels' reduction drain
for _ in range(3):
    ops.mlp_fused(y, res, w1, b1, w2, b2, gm, inplace=INPLACE)
torch.cuda.synchronize()
NR, NG, NE = 17, 64, 8
buf = torch.zeros(NR * NG * NE, dtype=torch.int64, device=dev)
L.check(L.lib().btsb_debug_mlp_trace(buf.data_ptr()), "trace on")
ops.mlp_fused(y, res, w1, b1, w2, b2, gm, inplace=INPLACE)
torch.cuda.synchronize()
L.check(L.lib().btsb_debug_mlp_trace(None), "trace off")
t = buf.cpu().view(NR, NG, NE)
t0 = int(t[0, 0, 0])
def rel(v):
    v = int(v)
    return "      -" if v == 0 else f"{v - t0:7d}"
NJ = 4 * C // 64
print(f"C={C} HW={HW} NJ={NJ}: clocks relative to the first G1 issue (block 0)")
torch.save(t, os.path.join("gpurun_out", f"mlp_trace_{C}.pt"))
print("  g tile j | G1rdy   G1iss   G2rdy   G2iss | ew  looptop  d1full  ld_done gelu_done hempty  hfull  d2epi_done")
for gidx in range(40):
    grp = gidx & 1
    # epilogue warps of group grp: k4 = (warp-2)>>2 with (k4 & 1) == grp -> ew in {4*grp .. 4*grp+3} u {8+4*grp ..}
    for ew in list(range(4 * grp, 4 * grp + 4)) + list(range(8 + 4 * grp, 12 + 4 * grp)):
        row = t[1 + ew, gidx]
        print(f"{gidx:3d} {gidx // NJ:4d} {gidx % NJ:2d} | {rel(t[0, gidx, 2])} {rel(t[0, gidx, 0])} {rel(t[0, gidx, 3])} {rel(t[0, gidx, 1])} | {ew:2d} " + " ".join(rel(v) for v in row[:7]))
