#!/usr/bin/env bash
OUT=gpurun_out/r02c2
mkdir -p $OUT
for v in auto plain; do
  if [ $v = auto ]; then envs="BTSB_X=0"; else envs="BTSB_HOST_PACK=0"; fi
  for rep in 1 2; do
  env $envs timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c2_$v$rep.log 2>$OUT/bench_c2_$v$rep.err; echo "bench c2 $v rc=$?"
  python - $OUT/bench_c2_$v$rep.log <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); e = d["e2e"]
        print("   value %.0f  e2e %.0f  ms/step %.3f  host_pack %s" % (d["value"], e["value"], e["ms_per_step"], e.get("host_pack")))
PY
  done
done
