mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_maxvit.py -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/t_mv.log 2>&1; echo "maxvit tests rc=$?"; tail -n 3 gpurun_out/t_mv.log
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.log 2>&1; echo "bench c4 rc=$?"
python scripts/show_bench.py gpurun_out/bench_c4.log > gpurun_out/bench_c4.txt; head -12 gpurun_out/bench_c4.txt | cut -c1-140
