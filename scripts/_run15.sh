mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu --tb=short -p no:cacheprovider -rA tests/test_gpu_maxvit.py > gpurun_out/t_maxvit.log 2>&1; echo "maxvit tests rc=$?"
grep -E "parity|passed|failed|FAILED|Error|error" gpurun_out/t_maxvit.log | head -40
timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 > gpurun_out/bench_c4.log 2>&1; echo "bench c4 rc=$?"
tail -c 600 gpurun_out/bench_c4.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_c4.csv python bench.py --workload c4 --batch 256 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
