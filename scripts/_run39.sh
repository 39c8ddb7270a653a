mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest gpu rc=$?"; tail -n 4 gpurun_out/t_all.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3.log 2>&1; echo "bench c3 rc=$?"
python scripts/show_bench.py gpurun_out/bench_c3.log > gpurun_out/bench_c3.txt; head -20 gpurun_out/bench_c3.txt
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.log 2>&1; echo "bench c4 rc=$?"
python scripts/show_bench.py gpurun_out/bench_c4.log > gpurun_out/bench_c4.txt; head -3 gpurun_out/bench_c4.txt | cut -c1-300
