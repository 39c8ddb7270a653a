#!/usr/bin/env bash
# Round-2 visit m: fp32-mode kernels (dw+LN v3 on fp32 rows, fp32 LN-patch, batched splitter loads), K1 ncu capture.
OUT=gpurun_out/r02m
mkdir -p $OUT
timeout 900 python -m pytest tests/ -q -m gpu -p no:cacheprovider -rA > $OUT/t_all.log 2>&1; echo "pytest gpu rc=$?"; tail -n 3 $OUT/t_all.log; grep -E "^(FAILED|ERROR)" $OUT/t_all.log | head
grep -h "^\[parity\]" $OUT/t_all.log > $OUT/parity_lines.txt; grep "fp32:" $OUT/parity_lines.txt | head -12
timeout 300 python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_fp32.log 2>$OUT/bench_fp32.err; echo "bench fp32 rc=$?"; tail -n 3 $OUT/bench_fp32.err
python scripts/show_bench.py $OUT/bench_fp32.log 2>/dev/null | cut -c1-170 | sed -n 1,22p
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --alerts 0"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cast_transpose63 -s 2 -c 1 -o $OUT/k1_cast -f $BENCH > $OUT/ncu_k1.log 2>&1; echo "ncu k1 rc=$?"
python scripts/ncu_summary.py $OUT/k1_cast.ncu-rep > $OUT/k1_cast.summary.txt 2>/dev/null; cat $OUT/k1_cast.summary.txt | cut -c1-500
