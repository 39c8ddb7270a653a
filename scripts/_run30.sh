mkdir -p gpurun_out
./scripts/micro/micro_fma2 | tee gpurun_out/micro_fma2.log
T="timeout 900 python -m pytest -q -m gpu --tb=short -p no:cacheprovider -x"
$T tests/test_gpu_kernels.py -k "dwln" > gpurun_out/t_dw.log 2>&1; echo "dwln tests rc=$?"; tail -n 3 gpurun_out/t_dw.log
timeout 300 python scripts/kbench.py --only "dwln" 2>&1 | grep dwln | tee gpurun_out/kbench_dw.log
$T tests/test_gpu_models.py > gpurun_out/t_models.log 2>&1; echo "model tests rc=$?"; tail -n 3 gpurun_out/t_models.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.log 2>&1; echo "bench c3 rc=$?"
python scripts/show_bench.py gpurun_out/bench_c3.log > gpurun_out/bench_c3.txt; head -12 gpurun_out/bench_c3.txt
