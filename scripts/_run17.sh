mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest gpu rc=$?"
tail -n 5 gpurun_out/t_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3.log 2>&1; echo "bench c3 rc=$?"
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/bench_c4.log 2>&1; echo "bench c4 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r01c_c3.csv $BENCH > gpurun_out/ncu_launches_c3.log 2>&1; echo "launch list c3 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r01c_c4.csv $BENCH --workload c4 --batch 1024 > gpurun_out/ncu_launches_c4.log 2>&1; echo "launch list c4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mlp_fused_kernel|dwln3_kernel|gemm_tc_kernel|lnpatch|stem|meta_head|im2col" -s 120 -c 60 -o gpurun_out/prof_r01c_c3 -f $BENCH > gpurun_out/ncu_full_c3.log 2>&1; echo "ncu full c3 rc=$?"
python scripts/show_bench.py gpurun_out/bench_c3.log gpurun_out/bench_c4.log 2>&1 | head -80
tail -c 600 gpurun_out/bench_ref.log
ls -la gpurun_out | tail -20
