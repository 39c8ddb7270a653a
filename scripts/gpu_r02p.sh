#!/usr/bin/env bash
# Round-2 visit p: polling waits restored, compile-time MLP tracing kept: kernel/model tests + C3 bench.
OUT=gpurun_out/r02p
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -q -m gpu -p no:cacheprovider > $OUT/t_km.log 2>&1; echo "pytest rc=$?"; tail -n 2 $OUT/t_km.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c3.log 2>$OUT/bench_c3.err; echo "bench rc=$?"
python scripts/show_bench.py $OUT/bench_c3.log 2>/dev/null | cut -c1-170 | sed -n 1,9p
