mkdir -p gpurun_out
BTSB_GEMM_2CTA=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gemm" -p no:cacheprovider --tb=short -rA > gpurun_out/t_g2.log 2>&1; echo "gemm tests (pair forced) rc=$?"; tail -n 3 gpurun_out/t_g2.log
grep -h "parity\] gemm bf16" gpurun_out/t_g2.log | head -12 | cut -c1-150
grep -h "Error\|error\|timed out" gpurun_out/t_g2.log | head -8
for m in 0 1; do echo "== BTSB_GEMM_2CTA=$m"; BTSB_GEMM_2CTA=$m timeout 300 python scripts/kbench.py --only "gemm" 2>&1 | grep gemm | tee -a gpurun_out/kbench_g2.log; done
