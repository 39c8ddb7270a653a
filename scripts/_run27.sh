mkdir -p gpurun_out
T="timeout 900 python -m pytest -q -m gpu --tb=short -p no:cacheprovider -x"
$T tests/test_gpu_kernels.py -k "gemm or stem" > gpurun_out/t_kern.log 2>&1; echo "kernel tests rc=$?"; tail -n 4 gpurun_out/t_kern.log
timeout 300 python scripts/kbench.py --only "gemm" 2>&1 | grep gemm | tee gpurun_out/kbench_gemm.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.log 2>&1; echo "bench c3 rc=$?"
python scripts/show_bench.py gpurun_out/bench_c3.log > gpurun_out/bench_c3.txt; head -8 gpurun_out/bench_c3.txt
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.log 2>&1; echo "bench c4 rc=$?"
python scripts/show_bench.py gpurun_out/bench_c4.log > gpurun_out/bench_c4.txt; head -3 gpurun_out/bench_c4.txt
