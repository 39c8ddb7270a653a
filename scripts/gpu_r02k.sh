#!/usr/bin/env bash
# Round-2 visit k (2 GPUs): c5 headline + default line with the collective-free replica measurement.
N=2
OUT=gpurun_out/r02k_n2
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 240 $TR bench.py --gpus $N --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_c5.log 2>$OUT/bench_c5.err; echo "bench c5 rc=$?"
grep -v "^\*\|OMP_NUM" $OUT/bench_c5.err | tail -n 12 | cut -c1-300
timeout 360 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_default.log 2> $OUT/bench_default.err; echo "bench default rc=$?"
python - $OUT <<'PY'
import json, sys
out = sys.argv[1]
for name in ("bench_c5", "bench_default"):
    for l in open(f"{out}/{name}.log"):
        if l.startswith("{"):
            d = json.loads(l)
            print(name, "n_gpus", d["n_gpus"], "value %.0f" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"])
            for k in ("c5", "collective"):
                if d.get(k) is not None:
                    print("   ", k, json.dumps(d[k])[:900])
PY
