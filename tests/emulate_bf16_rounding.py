"""CPU emulation of the bf16 mode's rounding points on the oracle (DESIGN.md lesson 20):
    python tests/emulate_bf16_rounding.py [case]      (case: a key of tests/cases.MODEL_CASES, default ff_pico)
Prints the max logit error against the fp32 oracle with each rounding (input, weights, residual stream, LayerNorm outputs,
hidden activations) switched on alone, all but one, all (eight draws with a 1e-6 jitter before every rounding), and with the
residual stream in bf16 / fp16 / fp32 (all stages or some).  Test infrastructure: imports oracle/."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))      # tests/ -> repo root
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from btsbot_b200 import synth
from oracle import convnext_oracle as O
import torch.nn.functional as F
from cases import MODEL_CASES, case_config
torch.set_num_threads(16)
gl = np.load(os.path.join(ROOT, "tests", "golden", "model_logits.npz"))
ex = np.load(os.path.join(ROOT, "tests", "golden", "example_inputs.npz"))
nsyn = int(gl["nsyn"])
trip = np.concatenate([ex["triplets"], synth.make_triplets(nsyn, start=1000)])
meta = np.concatenate([ex["metadata"], synth.make_metadata(nsyn, start=1000)])
img = torch.from_numpy(np.ascontiguousarray(trip.transpose(0, 3, 1, 2))); meta = torch.from_numpy(meta)
gen = torch.Generator().manual_seed(0)
def rb(x, on=True, jit=0.0):
    if not on: return x
    if jit: x = x * (1 + jit * torch.randn(x.shape, generator=gen))
    return x.bfloat16().float()
def trunk(sd, p, x, arch, R, jit):
    dims, depths = arch["dims"], arch["depths"]
    x = rb(x, R["in"], jit)
    x = F.conv2d(x, rb(sd[p + "stem.0.weight"], R["w"]), sd[p + "stem.0.bias"], stride=4)
    x = rb(O.layernorm2d(x, sd[p + "stem.1.weight"], sd[p + "stem.1.bias"]), R["res"], jit)
    for i, (c, d) in enumerate(zip(dims, depths)):
        if i > 0:
            q = f"{p}stages.{i}.downsample."
            x = rb(O.layernorm2d(x, sd[q + "0.weight"], sd[q + "0.bias"]), R["y"], jit)
            x = rb(F.conv2d(x, rb(sd[q + "1.weight"], R["w"]), sd[q + "1.bias"], stride=2), R["res"], jit)
        for j in range(d):
            q = f"{p}stages.{i}.blocks.{j}."
            s = x
            y = F.conv2d(x, sd[q + "conv_dw.weight"], sd[q + "conv_dw.bias"], padding=3, groups=c)
            y = rb(O.layernorm2d(y, sd[q + "norm.weight"], sd[q + "norm.bias"]), R["y"], jit)
            y = F.conv2d(y, rb(sd[q + "mlp.fc1.weight"], R["w"]), sd[q + "mlp.fc1.bias"])
            y = rb(F.gelu(y), R["h"], jit)
            y = F.conv2d(y, rb(sd[q + "mlp.fc2.weight"], R["w"]), sd[q + "mlp.fc2.bias"])
            x = rb(y * sd[q + "gamma"].view(1, -1, 1, 1) + s, R["res"], jit)
    return x
def run(case, R, jit=0.0):
    cfg = case_config(case)
    scale, shift = gl[case + "_cal"]
    sd = synth.to_torch(synth.apply_calibration(synth.make_state_dict(cfg, seed=2), cfg, 1.0, float(shift)))
    orig = O.trunk_features
    O.trunk_features = lambda sd_, p, x, arch, capture=None: trunk(sd_, p, x, arch, R, jit)
    try:
        with torch.no_grad():
            return O._forward(sd, cfg, img, meta, None).numpy()
    finally:
        O.trunk_features = orig
ALL = dict.fromkeys(["in", "w", "res", "y", "h"], True)
NONE = dict.fromkeys(ALL, False)
case = sys.argv[1] if len(sys.argv) > 1 else "ff_pico"
base = run(case, NONE)
print(case, "logit std", base.std())
for k in ALL:
    R = dict(NONE); R[k] = True
    print("only", k, "%.2e" % np.abs(run(case, R) - base).max())
for k in ALL:
    R = dict(ALL); R[k] = False
    print("all but", k, "%.2e" % np.abs(run(case, R) - base).max())
errs = [np.abs(run(case, ALL, jit=1e-6) - base).max() for _ in range(8)]
print("all, 8 draws:", ["%.2e" % e for e in errs])
print("---- residual-stream variants (all other roundings on) ----")
def rb16(x, jit=0.0):
    if jit: x = x * (1 + jit * torch.randn(x.shape, generator=gen))
    return x.half().float()
import types
def trunk2(sd, p, x, arch, mode, jit):
    dims, depths = arch["dims"], arch["depths"]
    def rr(x, stage):
        if mode == "fp16": return rb16(x, jit)
        if mode == "fp32": return x
        if mode == "fp32_s2" and stage == 2: return x
        if mode == "fp32_s12" and stage in (1, 2): return x
        return rb(x, True, jit)
    x = rb(x, True, jit)
    x = F.conv2d(x, rb(sd[p + "stem.0.weight"]), sd[p + "stem.0.bias"], stride=4)
    x = rr(O.layernorm2d(x, sd[p + "stem.1.weight"], sd[p + "stem.1.bias"]), 0)
    for i, (c, d) in enumerate(zip(dims, depths)):
        if i > 0:
            q = f"{p}stages.{i}.downsample."
            x = rb(O.layernorm2d(x, sd[q + "0.weight"], sd[q + "0.bias"]), True, jit)
            x = rr(F.conv2d(x, rb(sd[q + "1.weight"]), sd[q + "1.bias"], stride=2), i)
        for j in range(d):
            q = f"{p}stages.{i}.blocks.{j}."
            s = x
            y = F.conv2d(x, sd[q + "conv_dw.weight"], sd[q + "conv_dw.bias"], padding=3, groups=c)
            y = rb(O.layernorm2d(y, sd[q + "norm.weight"], sd[q + "norm.bias"]), True, jit)
            y = F.conv2d(y, rb(sd[q + "mlp.fc1.weight"]), sd[q + "mlp.fc1.bias"])
            y = rb(F.gelu(y), True, jit)
            y = F.conv2d(y, rb(sd[q + "mlp.fc2.weight"]), sd[q + "mlp.fc2.bias"])
            x = rr(y * sd[q + "gamma"].view(1, -1, 1, 1) + s, i)
    return x
def run2(case, mode, jit=0.0):
    cfg = case_config(case)
    scale, shift = gl[case + "_cal"]
    sd = synth.to_torch(synth.apply_calibration(synth.make_state_dict(cfg, seed=2), cfg, 1.0, float(shift)))
    orig = O.trunk_features
    O.trunk_features = lambda sd_, p, x, arch, capture=None: trunk2(sd_, p, x, arch, mode, jit)
    try:
        with torch.no_grad():
            return O._forward(sd, cfg, img, meta, None).numpy()
    finally:
        O.trunk_features = orig
for mode in ("bf16", "fp16", "fp32", "fp32_s2", "fp32_s12"):
    errs = [np.abs(run2(case, mode, jit=1e-6) - base).max() for _ in range(4)]
    print(mode, ["%.2e" % e for e in errs])
