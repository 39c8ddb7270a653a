mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -q -m gpu -p no:cacheprovider --tb=short -rA > gpurun_out/t_k.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/t_k.log
grep -h "parity" gpurun_out/t_k.log | grep -i "bf16" | tail -12 | cut -c1-220
timeout 300 python scripts/kbench.py --only "fc1+gelu" 2>&1 | grep gemm | tee gpurun_out/kbench_fc.log
timeout 300 python scripts/kbench.py --only mlp_fused 2>&1 | grep mlp_fused | tee gpurun_out/kbench_mlp.log
