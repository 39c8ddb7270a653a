mkdir -p gpurun_out
T="timeout 900 python -m pytest -q -m gpu --tb=short -rA -p no:cacheprovider"
$T tests/test_gpu_training.py > gpurun_out/t_train.log 2>&1; echo "train rc=$?"; grep -E "parity|passed|failed|Error|assert" gpurun_out/t_train.log | head -40
$T tests/test_gpu_kernels.py tests/test_gpu_models.py tests/test_gpu_preprocess.py > gpurun_out/t_rest.log 2>&1; echo "rest rc=$?"; tail -n 1 gpurun_out/t_rest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.log 2>&1; echo "bench bf16 rc=$?"
python scripts/show_bench.py gpurun_out/bench_bf16.log | head -8
