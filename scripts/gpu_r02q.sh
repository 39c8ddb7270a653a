#!/usr/bin/env bash
# Round-2 visit q: tf32 GEMM with a 4-slot TMA ring + 2-slot lo ring: fp32 parity + fp32 bench.
OUT=gpurun_out/r02q
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py tests/test_gpu_maxvit.py -q -m gpu -p no:cacheprovider -k "f32 or fp32 or graph" > $OUT/t_f32.log 2>&1; echo "pytest rc=$?"; tail -n 2 $OUT/t_f32.log; grep -E "^(FAILED|ERROR)" $OUT/t_f32.log | head -5
timeout 300 python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_fp32.log 2>$OUT/bench_fp32.err; echo "bench fp32 rc=$?"; tail -n 3 $OUT/bench_fp32.err
python scripts/show_bench.py $OUT/bench_fp32.log 2>/dev/null | cut -c1-170 | sed -n 1,14p
