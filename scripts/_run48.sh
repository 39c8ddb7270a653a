mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "mlp_fused" -p no:cacheprovider --tb=short -rA > gpurun_out/t_mlp.log 2>&1; echo "mlp tests rc=$?"; tail -n 3 gpurun_out/t_mlp.log
grep -h "parity\] mlp_fused\|Error\|timed out\|stalled" gpurun_out/t_mlp.log | head -14 | cut -c1-150
timeout 300 python scripts/kbench.py --only mlp_fused 2>&1 | grep mlp_fused | tee gpurun_out/kbench_mlp.log
timeout 300 python scripts/kbench.py --only "C=320" 2>&1 | grep gemm | tee -a gpurun_out/kbench_mlp.log
