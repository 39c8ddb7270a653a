#!/usr/bin/env bash
# Round-2 visit r: wide fused MLP in place (bulk tensor reduction drain): parity + A/B.
OUT=gpurun_out/r02r
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "mlp_fused" -rA > $OUT/t_k.log 2>&1; echo "pytest kernels rc=$?"; tail -n 2 $OUT/t_k.log; grep "in place" $OUT/t_k.log | head
BTSB_MLP_INPLACE=1 timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_maxvit.py -q -m gpu -p no:cacheprovider -rA > $OUT/t_m.log 2>&1; echo "pytest models (in place) rc=$?"; tail -n 2 $OUT/t_m.log; grep -E "^(FAILED|ERROR)" $OUT/t_m.log | head; grep "\[parity\].*bf16: gain" $OUT/t_m.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c3.log 2>$OUT/bench_c3.err; echo "bench rc=$?"
BTSB_MLP_INPLACE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c3_inplace.log 2>$OUT/bench_c3_inplace.err; echo "bench in place rc=$?"; tail -n 3 $OUT/bench_c3_inplace.err
for f in bench_c3 bench_c3_inplace; do python scripts/show_bench.py $OUT/$f.log 2>/dev/null | cut -c1-170 | sed -n 1,6p; done
BTSB_MLP_INPLACE=1 timeout 90 python scripts/mlp_trace.py 320 9 > $OUT/mlp_trace_320.txt 2>&1
