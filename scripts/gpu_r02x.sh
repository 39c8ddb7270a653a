#!/usr/bin/env bash
# Round-2 visit x: dedicated D2-epilogue warps in the narrow fused MLP (C <= 160) + dwln5 at the pico widths: parity + bench + traces.
OUT=gpurun_out/r02x
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -rA -x -k "mlp_fused or dwln" > $OUT/t_k.log 2>&1; echo "pytest kernels rc=$?"; tail -n 2 $OUT/t_k.log; grep -E "^(FAILED|ERROR)" $OUT/t_k.log | head
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_maxvit.py -q -m gpu -p no:cacheprovider -rA > $OUT/t_m.log 2>&1; echo "pytest models rc=$?"; tail -n 2 $OUT/t_m.log; grep -E "^(FAILED|ERROR)" $OUT/t_m.log | head; grep "\[parity\].*bf16: gain" $OUT/t_m.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c3.log 2>$OUT/bench_c3.err; echo "bench rc=$?"; tail -n 2 $OUT/bench_c3.err
python scripts/show_bench.py $OUT/bench_c3.log 2>/dev/null | cut -c1-170 | sed -n 1,17p
timeout 90 python scripts/mlp_trace.py 80 225 > $OUT/mlp_trace_80.txt 2>&1
timeout 90 python scripts/mlp_trace.py 160 49 > $OUT/mlp_trace_160.txt 2>&1
grep -E "\|  ?(0|4)  " $OUT/mlp_trace_80.txt | sed -n 1,24p | cut -c1-130
