mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -q -m gpu -p no:cacheprovider --tb=line -rA > gpurun_out/t_k.log 2>&1; echo "tests v4 rc=$?"; tail -n 2 gpurun_out/t_k.log
grep -h "parity\] .* bf16: gain-1" gpurun_out/t_k.log | cut -c1-120
BTSB_DWLN=3 timeout 900 python -m pytest tests/test_gpu_models.py -q -m gpu -p no:cacheprovider --tb=line -rA -k "bf16_logits" > gpurun_out/t_k3.log 2>&1; echo "tests v3 rc=$?"; tail -n 2 gpurun_out/t_k3.log
grep -h "parity\] .* bf16: gain-1" gpurun_out/t_k3.log | cut -c1-120
grep -h "assert np.float32" gpurun_out/t_k.log gpurun_out/t_k3.log
timeout 300 python scripts/kbench.py --only "dwln" 2>&1 | grep dwln | tee gpurun_out/kbench_dw.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.log 2>&1; echo "bench c3 rc=$?"
python scripts/show_bench.py gpurun_out/bench_c3.log > gpurun_out/bench_c3.txt; head -9 gpurun_out/bench_c3.txt | cut -c1-160
