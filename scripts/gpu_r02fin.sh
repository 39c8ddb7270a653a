#!/usr/bin/env bash
# final visit: the default bench line exactly as the driver launches it (no flags), wall-clocked; then the reference arm.
OUT=gpurun_out/r02fin
mkdir -p $OUT
SECONDS=0; timeout 900 python bench.py > $OUT/bench_default.log 2>$OUT/bench_default.err; echo "bench default rc=$?"; echo "wall seconds: $SECONDS"
python - $OUT/bench_default.log <<'PY'
import json, sys
n = 0
for l in open(sys.argv[1]):
    if l.startswith("{"):
        n += 1
        d = json.loads(l); e = d["e2e"]
        print("   value %.0f  ms/step %.3f  e2e %.0f  (%.3f ms)  launches %d  host_pack %s" % (d["value"], d["ms_per_step"], e["value"], e["ms_per_step"], d["gpu_launches"], e.get("host_pack")))
        for k in ("c5", "c4", "c2", "fp32"):
            if d.get(k) is not None: print("   ", k, "value %.0f  e2e %.0f  %s" % (d[k]["value"], d[k]["e2e"]["value"], d[k]["e2e"].get("host_pack")))
        print("    roofline", {k: d["roofline"][k] for k in ("kernel", "bound", "achieved", "peak", "frac", "frac_burst", "traffic")})
        print("    cpu_baseline", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["kind"])
        print("    clocks", d["clocks"])
print("json lines:", n)
PY
