#!/usr/bin/env bash
# Round-2 visit w (gpurun --gpus N): e2e scaling with and without host-side bf16 packing at N ranks, then the full default
# line exactly as the driver launches it (C3 + c5 / c4 / c2 / fp32 sub-records).
N=${1:-4}
OUT=gpurun_out/r02w_n$N
mkdir -p $OUT
nproc > $OUT/host.txt; lscpu | grep -E "Model name|Socket|Core|NUMA" >> $OUT/host.txt
nvidia-smi topo -m > $OUT/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
show() {
python - $1 <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); e = d["e2e"]
        print("   n_gpus", d["n_gpus"], "value %.0f  e2e %.0f  ms/step %.3f  h2d/gpu %.1f GB/s  pcie %s  host_pack %s" % (d["value"], e["value"], e["ms_per_step"], e["h2d_gbs_per_gpu"], e.get("pcie", {}).get("h2d_gbs_per_gpu_bare"), e.get("host_pack")))
        for k in ("c5", "c4", "c2", "fp32"):
            if d.get(k) is not None: print("   ", k, json.dumps(d[k])[:400])
PY
}
for v in plain auto packed; do
  case $v in
    auto) envs="BTSB_X=0";;
    plain) envs="BTSB_HOST_PACK=0";;
    packed) envs="BTSB_HOST_PACK=1";;
  esac
  env $envs timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c3_$v.log 2>$OUT/bench_c3_$v.err; echo "bench $v rc=$?"; tail -n 2 $OUT/bench_c3_$v.err | cut -c1-200
  show $OUT/bench_c3_$v.log
done
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_default.log 2>$OUT/bench_default.err; echo "bench default rc=$?"; tail -n 3 $OUT/bench_default.err | cut -c1-200
show $OUT/bench_default.log
timeout 240 $TR tests/ddp_check.py > $OUT/ddp_check.log 2>&1; echo "ddp_check rc=$?"; tail -n 2 $OUT/ddp_check.log
