mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r02.csv $BENCH > gpurun_out/ncu_launches_r02.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_fused_kernel -s 12 -c 3 -o gpurun_out/prof_r02_fused -f $BENCH > gpurun_out/ncu_fused.log 2>&1; echo "ncu fused rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 72 -c 5 -o gpurun_out/prof_r02_gemm -f $BENCH > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dwln3_kernel -s 42 -c 5 -o gpurun_out/prof_r02_dwln3 -f $BENCH > gpurun_out/ncu_dwln3.log 2>&1; echo "ncu dwln3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lnpatch|meta_head" -s 12 -c 4 -o gpurun_out/prof_r02_misc -f $BENCH > gpurun_out/ncu_misc.log 2>&1; echo "ncu misc rc=$?"
ls -la gpurun_out | tail
