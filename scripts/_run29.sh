for d in 0 1 4 8; do echo "== BTSB_GEMM_DBG=$d"; BTSB_GEMM_DBG=$d timeout 300 python scripts/kbench.py --only "C=320" 2>&1 | grep gemm; done | tee gpurun_out/kbench_gemm_dbg2.log
