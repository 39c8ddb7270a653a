mkdir -p gpurun_out
T="timeout 900 python -m pytest -q -m gpu --tb=short -p no:cacheprovider -x"
$T tests/test_gpu_kernels.py -k "mlp_fused" > gpurun_out/t_mlp.log 2>&1; echo "mlp tests rc=$?"; tail -n 15 gpurun_out/t_mlp.log
$T tests/test_gpu_kernels.py -k "not mlp_fused" > gpurun_out/t_kern.log 2>&1; echo "kernel tests rc=$?"; tail -n 5 gpurun_out/t_kern.log
$T tests/test_gpu_models.py > gpurun_out/t_models.log 2>&1; echo "model tests rc=$?"; tail -n 5 gpurun_out/t_models.log
timeout 300 python scripts/kbench.py > gpurun_out/kbench_v2.log 2>&1; echo "kbench rc=$?"; cat gpurun_out/kbench_v2.log
BTSB_MLP_V1=1 timeout 300 python scripts/kbench.py --only mlp_fused > gpurun_out/kbench_v1.log 2>&1; cat gpurun_out/kbench_v1.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.log 2>&1; echo "bench c3 rc=$?"
python scripts/show_bench.py gpurun_out/bench_c3.log | head -24
