mkdir -p gpurun_out
T="timeout 900 python -m pytest -q -m gpu --tb=short -p no:cacheprovider -x"
$T tests/test_gpu_kernels.py -k "mlp_fused or gemm" > gpurun_out/t_kern.log 2>&1; echo "kernel tests rc=$?"; tail -n 4 gpurun_out/t_kern.log
timeout 300 python scripts/kbench.py --only "mlp_fused" 2>&1 | tee gpurun_out/kbench_mlp.log
for d in 0 1 2 3; do echo "== BTSB_GEMM_DBG=$d"; BTSB_GEMM_DBG=$d timeout 300 python scripts/kbench.py --only "gemm" 2>&1 | grep gemm; done | tee gpurun_out/kbench_gemm_dbg.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.log 2>&1; echo "bench c3 rc=$?"
python scripts/show_bench.py gpurun_out/bench_c3.log | head -12
