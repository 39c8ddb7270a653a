mkdir -p gpurun_out
T="timeout 600 python -m pytest -q -m gpu --tb=short -p no:cacheprovider -x"
for e in 1 0; do BTSB_MLP_HSMEM=$e $T tests/test_gpu_kernels.py -k "mlp_fused" > gpurun_out/t_mlp_$e.log 2>&1; echo "mlp tests HSMEM=$e rc=$?"; tail -n 6 gpurun_out/t_mlp_$e.log; done
for e in 1 0; do echo "== BTSB_MLP_HSMEM=$e"; BTSB_MLP_HSMEM=$e timeout 300 python scripts/kbench.py --only "mlp_fused" 2>&1 | grep mlp_fused; done | tee gpurun_out/kbench_mlp_ht.log
$T tests/test_gpu_models.py > gpurun_out/t_models.log 2>&1; echo "model tests rc=$?"; tail -n 3 gpurun_out/t_models.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.log 2>&1; echo "bench c3 rc=$?"
python scripts/show_bench.py gpurun_out/bench_c3.log > gpurun_out/bench_c3.txt; head -8 gpurun_out/bench_c3.txt
