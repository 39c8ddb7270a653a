#!/usr/bin/env bash
# Multi-GPU visit (gpurun --gpus N -- 'bash scripts/gpu_multi.sh N <tag>'): DDP == large-batch check at world N, the default
# bench line (C3 + c5 / c4 / c2 / fp32 sub-records, launched exactly as the driver does) and the C5 headline line.
N=${1:-2}
TAG=${2:-multi_n$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT/gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 240 $TR tests/ddp_check.py > $OUT/ddp_check.log 2>&1; echo "ddp_check rc=$?"; tail -n 2 $OUT/ddp_check.log
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 360 $TR bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_default.log 2> $OUT/bench_default.err; echo "bench default rc=$?"
grep -E "NVLS|Ring|Tree|nranks|comm 0x" $OUT/bench_default.log $OUT/bench_default.err | grep -m 12 -E "NVLS|Connected|Channel 00|nranks" | cut -c1-200 > $OUT/nccl_lines.txt
timeout 240 $TR bench.py --gpus $N --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_c5.log 2>$OUT/bench_c5.err; echo "bench c5 rc=$?"
python - $OUT <<'PY'
import json, sys
out = sys.argv[1]
for name in ("bench_default", "bench_c5"):
    for l in open(f"{out}/{name}.log"):
        if l.startswith("{"):
            d = json.loads(l)
            print(name, "n_gpus", d["n_gpus"], "value %.0f" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"])
            for k in ("c5", "c4", "c2", "fp32", "sustained", "collective"):
                if d.get(k) is not None:
                    print("   ", k, json.dumps(d[k])[:600])
            if "pcie" in d["e2e"]:
                print("    pcie", d["e2e"]["pcie"], "h2d_gbs_per_gpu", d["e2e"]["h2d_gbs_per_gpu"])
PY
tail -n 5 $OUT/bench_default.err | cut -c1-300
