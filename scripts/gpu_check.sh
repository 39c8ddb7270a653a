#!/usr/bin/env bash
# One GPU-box visit: per-kernel parity (SIMT and tcgen05 in separate processes), model parity, smoke, bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
T="timeout 600 python -m pytest -q -m gpu --tb=short -rA -p no:cacheprovider"
$T tests/test_gpu_preprocess.py > gpurun_out/t_pre.log 2>&1; echo "pre rc=$?"
$T tests/test_gpu_kernels.py -k "not tcgen05" > gpurun_out/t_simt.log 2>&1; echo "simt rc=$?"
$T tests/test_gpu_kernels.py -k "tcgen05" > gpurun_out/t_tc.log 2>&1; echo "tc rc=$?"
$T tests/test_gpu_models.py -k "fp32" > gpurun_out/t_models_fp32.log 2>&1; echo "models fp32 rc=$?"
$T tests/test_gpu_models.py -k "not fp32" > gpurun_out/t_models_rest.log 2>&1; echo "models rest rc=$?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --precision fp32 --batch 2048 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32.log 2>&1; echo "bench fp32 rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_bf16.log 2>&1; echo "bench bf16 rc=$?"
for f in gpurun_out/t_*.log; do echo "== $f"; tail -n 3 $f; done
tail -n 5 gpurun_out/smoke.log
tail -c 1500 gpurun_out/bench_bf16.log
