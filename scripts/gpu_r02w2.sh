#!/usr/bin/env bash
# Round-2 visit w2 (gpurun --gpus N): e2e at N ranks with the split host packing (auto) against the plain copy and two fixed fractions.
N=${1:-4}
OUT=gpurun_out/r02w2_n$N
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
show() {
python - $1 <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); e = d["e2e"]
        print("   n_gpus", d["n_gpus"], "value %.0f  e2e %.0f  ms/step %.3f  h2d/gpu %.1f GB/s  host_pack %s" % (d["value"], e["value"], e["ms_per_step"], e["h2d_gbs_per_gpu"], e.get("host_pack")))
PY
}
for v in auto 0 0.3 0.5; do
  if [ $v = auto ]; then envs="BTSB_X=0"; else envs="BTSB_HOST_PACK=$v"; fi
  env $envs timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c3_$v.log 2>$OUT/bench_c3_$v.err; echo "bench $v rc=$?"; tail -n 1 $OUT/bench_c3_$v.err | cut -c1-200
  show $OUT/bench_c3_$v.log
done
