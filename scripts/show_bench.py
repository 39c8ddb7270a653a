#!/usr/bin/env python
"""Pretty-print bench.py JSON lines: python scripts/show_bench.py gpurun_out/bench_bf16.log"""
import json
import sys

for f in sys.argv[1:]:
    for line in open(f):
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        if d.get("impl") == "reference":
            print(f, "REFERENCE %.0f alerts/s" % d["value"])
            continue
        print(f, "value %.0f alerts/s  ms/step %.3f  e2e %.0f  launches %d  n_gpus %d clocks %s" % (
            d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["n_gpus"], d["clocks"]))
        print("  roofline", d["roofline"])
        for k, v in d["kernels"].items():
            roof = "  %.2f of %s peak" % (v["frac"], v["bound"]) if "frac" in v else ""
            print("   %-16s n=%4.1f  %8.3f ms/launch  %7.0f GB/s %7.1f TF/s  share %.3f%s" % (
                k, v["launches_per_step"], v["ms_per_launch"], v["gbs"], v["tflops"], v["share"], roof))
        print("  cpu", d["cpu_baseline"])
