mkdir -p gpurun_out/r01i
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01i/launches_c3.csv $BENCH > gpurun_out/r01i/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_fused2_kernel -s 6 -c 1 -o gpurun_out/r01i/mlp2_80 -f $BENCH > gpurun_out/r01i/ncu_mlp.log 2>&1; echo "ncu mlp rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwln3_kernel -s 6 -c 1 -o gpurun_out/r01i/dwln15 -f $BENCH > gpurun_out/r01i/ncu_dw.log 2>&1; echo "ncu dwln rc=$?"
ls -la gpurun_out/r01i
