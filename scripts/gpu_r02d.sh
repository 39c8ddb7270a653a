#!/usr/bin/env bash
# Round-2 visit d: new tests (Scorer invalidation, ff_maxvit, LS training, dwln5 determinism), K1 fast path, FMA micro-benchmark.
TAG=${1:-r02d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 60 scripts/micro/micro_fma2 > $OUT/micro_fma2.txt 2>&1; echo "micro_fma2 rc=$?"; cat $OUT/micro_fma2.txt
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider -rA > $OUT/t_all.log 2>&1; echo "pytest gpu rc=$?"; tail -n 3 $OUT/t_all.log; grep -E "^(FAILED|ERROR)" $OUT/t_all.log | head
grep -h "^\[parity\]" $OUT/t_all.log > $OUT/parity_lines.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_c3.log 2>$OUT/bench_c3.err; echo "bench rc=$?"; tail -n 3 $OUT/bench_c3.err
python scripts/show_bench.py $OUT/bench_c3.log 2>/dev/null | cut -c1-170 | head -24
