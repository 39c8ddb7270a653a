#!/usr/bin/env bash
# Round-2 visit z: split host packing (fraction f of each batch rounded to bf16 on the host beside the fp32 DMA): parity + e2e A/B.
OUT=gpurun_out/r02z
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_preprocess.py -q -m gpu -p no:cacheprovider -rA -k "alert_scorer or host_pack or preprocess or cast" > $OUT/t_pack.log 2>&1; echo "pytest pack rc=$?"; tail -n 2 $OUT/t_pack.log; grep -E "^(FAILED|ERROR)" $OUT/t_pack.log | head
show() {
python - $1 <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); e = d["e2e"]
        print("   value %.0f  e2e %.0f  ms/step %.3f  h2d %.1f GB/s  host_pack %s" % (d["value"], e["value"], e["ms_per_step"], e["h2d_gbs_per_gpu"], e.get("host_pack")))
PY
}
for v in auto 0 0.35 0.5 0.65; do
  if [ $v = auto ]; then envs="BTSB_X=0"; else envs="BTSB_HOST_PACK=$v"; fi
  env $envs timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c3_$v.log 2>$OUT/bench_c3_$v.err; echo "bench $v rc=$?"; tail -n 2 $OUT/bench_c3_$v.err
  show $OUT/bench_c3_$v.log
done
BTSB_BENCH_C3_GRAPH=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c3_graph_auto.log 2>$OUT/bench_c3_graph_auto.err; echo "bench graph auto rc=$?"; show $OUT/bench_c3_graph_auto.log
BTSB_BENCH_C3_GRAPH=1 BTSB_HOST_PACK=0.6 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c3_graph_06.log 2>$OUT/bench_c3_graph_06.err; echo "bench graph 0.6 rc=$?"; show $OUT/bench_c3_graph_06.log
