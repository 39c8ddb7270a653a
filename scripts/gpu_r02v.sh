#!/usr/bin/env bash
# Round-2 visit v: host-side bf16 packing in the e2e path (AlertScorer host_pack): parity + e2e A/B (plain / packed / packed + graph).
OUT=gpurun_out/r02v
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_preprocess.py -q -m gpu -p no:cacheprovider -rA -k "alert_scorer or host_pack or preprocess or cast" > $OUT/t_pack.log 2>&1; echo "pytest pack rc=$?"; tail -n 2 $OUT/t_pack.log; grep -E "^(FAILED|ERROR)" $OUT/t_pack.log | head
for v in auto plain packed; do
  case $v in
    auto) envs="BTSB_X=0";;
    plain) envs="BTSB_HOST_PACK=0";;
    packed) envs="BTSB_HOST_PACK=1";;
  esac
  env $envs timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c3_$v.log 2>$OUT/bench_c3_$v.err; echo "bench $v rc=$?"; tail -n 2 $OUT/bench_c3_$v.err
  python - $OUT/bench_c3_$v.log <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); e = d["e2e"]
        print("   value %.0f  e2e %.0f  ms/step %.3f  h2d %.1f GB/s  host_pack %s" % (d["value"], e["value"], e["ms_per_step"], e["h2d_gbs_per_gpu"], e.get("host_pack")))
PY
done
BTSB_BENCH_C3_GRAPH=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c3_graph.log 2>$OUT/bench_c3_graph.err; echo "bench graph rc=$?"; tail -n 2 $OUT/bench_c3_graph.err
python - $OUT/bench_c3_graph.log <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); e = d["e2e"]
        print("   value %.0f  e2e %.0f  ms/step %.3f  h2d %.1f GB/s  host_pack %s" % (d["value"], e["value"], e["ms_per_step"], e["h2d_gbs_per_gpu"], e.get("host_pack")))
PY
timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c2.log 2>$OUT/bench_c2.err; echo "bench c2 rc=$?"; python scripts/show_bench.py $OUT/bench_c2.log 2>/dev/null | cut -c1-170 | sed -n 1,2p
