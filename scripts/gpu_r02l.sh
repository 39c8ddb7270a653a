#!/usr/bin/env bash
# Round-2 visit l: inference CUDA graph (tests + default bench line with the graphed C2 sub-record), then ncu evidence.
OUT=gpurun_out/r02l
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_maxvit.py tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider > $OUT/t_m.log 2>&1; echo "pytest rc=$?"; tail -n 3 $OUT/t_m.log; grep -E "^(FAILED|ERROR)" $OUT/t_m.log | head
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_default.log 2> $OUT/bench_default.err; echo "bench default rc=$?"; tail -n 3 $OUT/bench_default.err
BTSB_BENCH_GRAPH=0 timeout 200 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c2_eager.log 2>$OUT/bench_c2_eager.err; echo "bench c2 eager rc=$?"
timeout 200 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c2_graph.log 2>$OUT/bench_c2_graph.err; echo "bench c2 graph rc=$?"; tail -n 3 $OUT/bench_c2_graph.err
python - <<'PY'
import json
for name in ("bench_default", "bench_c2_eager", "bench_c2_graph"):
    for l in open(f"gpurun_out/r02l/{name}.log"):
        if l.startswith("{"):
            d = json.loads(l)
            print(name, "value %.0f" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "launches", d["gpu_launches"])
            for k in ("c5", "c4", "c2", "fp32"):
                if d.get(k): print("   ", k, "value %.0f" % d[k]["value"], "ms/step %.3f" % d[k]["ms_per_step"], "e2e %.0f" % d[k]["e2e"]["value"], "launches", d[k]["gpu_launches"])
PY
bash scripts/gpu_profile.sh r02l
