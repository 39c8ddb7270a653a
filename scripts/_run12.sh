mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --batch 4096 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_fused_kernel -s 4 -c 1 -o gpurun_out/prof_r01d_fused -f $BENCH > gpurun_out/ncu_fused.log 2>&1; echo "ncu fused rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 46 -c 2 -o gpurun_out/prof_r01d_gemm -f $BENCH > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwln3_kernel -s 4 -c 1 -o gpurun_out/prof_r01d_dwln3 -f $BENCH > gpurun_out/ncu_dwln3.log 2>&1; echo "ncu dwln3 rc=$?"
