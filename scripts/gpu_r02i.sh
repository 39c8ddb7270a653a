#!/usr/bin/env bash
# Round-2 visit i: full suite (tcgen05 attention, K1 double-precision quotient, all-alert label checks), attention A/B on C4.
TAG=${1:-r02i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/ -q -m gpu -p no:cacheprovider -rA > $OUT/t_all.log 2>&1; echo "pytest gpu rc=$?"; tail -n 3 $OUT/t_all.log; grep -E "^(FAILED|ERROR)" $OUT/t_all.log | head
grep -h "^\[parity\]" $OUT/t_all.log > $OUT/parity_lines.txt
grep -h "labels differ\|attention\|crop s=" $OUT/parity_lines.txt | cut -c1-200 | head -40
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_c4.log 2>$OUT/bench_c4.err; echo "bench c4 rc=$?"; tail -n 3 $OUT/bench_c4.err
BTSB_ATTN_MMA=1 timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_c4_mma.log 2>$OUT/bench_c4_mma.err; echo "bench c4 mma rc=$?"
for f in bench_c4 bench_c4_mma; do python scripts/show_bench.py $OUT/$f.log 2>/dev/null | grep -E "value|attn" | cut -c1-170; done
timeout 120 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 $OUT/smoke.log
