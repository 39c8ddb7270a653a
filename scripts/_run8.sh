mkdir -p gpurun_out
T="timeout 900 python -m pytest -q -m gpu --tb=short -rA -p no:cacheprovider"
$T tests/test_gpu_kernels.py -k "tcgen05" > gpurun_out/t_tc.log 2>&1; echo "tc rc=$?"; tail -n 1 gpurun_out/t_tc.log
$T tests/test_gpu_models.py > gpurun_out/t_models.log 2>&1; echo "models rc=$?"; tail -n 1 gpurun_out/t_models.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.log 2>&1; echo "bench bf16 rc=$?"
python scripts/show_bench.py gpurun_out/bench_bf16.log > gpurun_out/bench_bf16.txt 2>&1; head -20 gpurun_out/bench_bf16.txt
BENCH="python bench.py --steps 1 --warmup 3 --batch 4096 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwln2_kernel -s 14 -c 1 -o gpurun_out/prof_r01c_dwln15 -f $BENCH > gpurun_out/ncu_dwln15.log 2>&1; echo "ncu dwln15 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_fused_kernel -s 4 -c 1 -o gpurun_out/prof_r01c_fused -f $BENCH > gpurun_out/ncu_fused.log 2>&1; echo "ncu fused rc=$?"
