#!/usr/bin/env bash
# Round-2 visit c: full GPU suite after the cleanup + new bench.py line with sub-records.
TAG=${1:-r02c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider -rA > $OUT/t_all.log 2>&1; echo "pytest gpu rc=$?"; tail -n 3 $OUT/t_all.log
grep -h "^\[parity\]" $OUT/t_all.log > $OUT/parity_lines.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_default.log 2> $OUT/bench_default.err; echo "bench default rc=$?"; tail -n 5 $OUT/bench_default.err
python - <<'PY'
import json,sys
for l in open("gpurun_out/r02c/bench_default.log"):
    if l.startswith("{"):
        d=json.loads(l)
        print("value",d["value"],"e2e",d["e2e"])
        for k in ("c5","c4","c2","fp32","sustained","gpu_library_baseline"):
            print(k, json.dumps(d.get(k))[:700])
        print("roofline", d["roofline"])
        for k,v in d["kernels"].items(): print("  ",k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
        print("cpu", d["cpu_baseline"])
PY
