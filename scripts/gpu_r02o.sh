#!/usr/bin/env bash
# Round-2 visit o: mbarrier waits parked with a suspend-time hint, compile-time MLP tracing: full suite + C3/C4/C5 benches.
OUT=gpurun_out/r02o
mkdir -p $OUT
timeout 900 python -m pytest tests/ -q -m gpu -p no:cacheprovider > $OUT/t_all.log 2>&1; echo "pytest gpu rc=$?"; tail -n 3 $OUT/t_all.log; grep -E "^(FAILED|ERROR)" $OUT/t_all.log | head
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_default.log 2>$OUT/bench_default.err; echo "bench rc=$?"; tail -n 3 $OUT/bench_default.err
python scripts/show_bench.py $OUT/bench_default.log 2>/dev/null | cut -c1-170 | sed -n 1,20p
python - <<'PY'
import json
for l in open("gpurun_out/r02o/bench_default.log"):
    if l.startswith("{"):
        d = json.loads(l)
        for k in ("c5", "c4", "c2", "fp32"):
            if d.get(k): print("   ", k, "value %.0f" % d[k]["value"], "ms/step %.3f" % d[k]["ms_per_step"], "e2e %.0f" % d[k]["e2e"]["value"])
PY
timeout 120 python scripts/mlp_trace.py 80 225 > $OUT/mlp_trace_80.txt 2>&1; echo "trace rc=$?"; head -n 4 $OUT/mlp_trace_80.txt | cut -c1-150
