mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -q -m gpu -p no:cacheprovider --tb=short -x > gpurun_out/t_k.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/t_k.log
timeout 300 python scripts/kbench.py --only "dwln" 2>&1 | grep dwln | tee gpurun_out/kbench_dw.log
BTSB_DWLN=3 timeout 300 python scripts/kbench.py --only "dwln" 2>&1 | grep dwln | sed 's/^/v3: /' | tee -a gpurun_out/kbench_dw.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.log 2>&1; echo "bench c3 rc=$?"
python scripts/show_bench.py gpurun_out/bench_c3.log > gpurun_out/bench_c3.txt; head -12 gpurun_out/bench_c3.txt | cut -c1-160
