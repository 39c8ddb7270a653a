#!/usr/bin/env bash
# N-GPU visit (gpurun --gpus N -- 'bash scripts/gpu_n8.sh N'): the default bench line exactly as the driver launches it.
N=${1:-8}
OUT=gpurun_out/r02n$N
mkdir -p $OUT
nproc > $OUT/host.txt; lscpu | grep -E "Model name|Socket|Core|NUMA" >> $OUT/host.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_default.log 2>$OUT/bench_default.err; echo "bench default rc=$?"; tail -n 3 $OUT/bench_default.err | cut -c1-200
python - $OUT/bench_default.log <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); e = d["e2e"]
        print("   n_gpus", d["n_gpus"], "value %.0f  e2e %.0f  ms/step %.3f  h2d/gpu %.1f GB/s  host_pack %s" % (d["value"], e["value"], e["ms_per_step"], e["h2d_gbs_per_gpu"], e.get("host_pack")))
        for k in ("c5", "c4", "c2", "fp32"):
            if d.get(k) is not None: print("   ", k, json.dumps({kk: vv for kk, vv in d[k].items() if kk not in ("workload", "kernels", "clocks")})[:500])
PY
timeout 240 $TR tests/ddp_check.py > $OUT/ddp_check.log 2>&1; echo "ddp_check rc=$?"; tail -n 2 $OUT/ddp_check.log
