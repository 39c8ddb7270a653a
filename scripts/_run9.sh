mkdir -p gpurun_out
T="timeout 900 python -m pytest -q -m gpu --tb=short -rA -p no:cacheprovider"
$T tests/test_gpu_kernels.py -k "tcgen05 or dwln or lnpatch" > gpurun_out/t_tc.log 2>&1; echo "tc rc=$?"; tail -n 1 gpurun_out/t_tc.log; grep -E "^FAILED|Error" gpurun_out/t_tc.log | head
$T tests/test_gpu_models.py > gpurun_out/t_models.log 2>&1; echo "models rc=$?"; tail -n 1 gpurun_out/t_models.log; grep -E "^FAILED" gpurun_out/t_models.log | head
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.log 2>&1; echo "bench bf16 rc=$?"
python scripts/show_bench.py gpurun_out/bench_bf16.log > gpurun_out/bench_bf16.txt 2>&1; head -22 gpurun_out/bench_bf16.txt
