#!/usr/bin/env bash
# Round-2 visit z2: packed-fraction sweep of the e2e path.
OUT=gpurun_out/r02z2
mkdir -p $OUT
show() {
python - $1 <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); e = d["e2e"]
        print("   value %.0f  e2e %.0f  ms/step %.3f  h2d %.1f GB/s  host_pack %s" % (d["value"], e["value"], e["ms_per_step"], e["h2d_gbs_per_gpu"], e.get("host_pack")))
PY
}
for v in 0.7 0.8 0.9 1.0; do
  BTSB_HOST_PACK=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c3_$v.log 2>$OUT/bench_c3_$v.err; echo "bench $v rc=$?"; tail -n 1 $OUT/bench_c3_$v.err
  show $OUT/bench_c3_$v.log
done
for v in 0.75 0.9; do
  BTSB_BENCH_C3_GRAPH=1 BTSB_HOST_PACK=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c3_graph_$v.log 2>$OUT/bench_c3_graph_$v.err; echo "bench graph $v rc=$?"; show $OUT/bench_c3_graph_$v.log
done
