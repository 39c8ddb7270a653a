mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -q -m gpu -p no:cacheprovider -rA --tb=short > gpurun_out/t_train.log 2>&1; echo "pytest training rc=$?"
grep -E "parity|passed|failed|Error|error" gpurun_out/t_train.log | tail -n 25
timeout 600 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5.log 2>&1; echo "bench c5 rc=$?"
python scripts/show_bench.py gpurun_out/bench_c5.log > gpurun_out/bench_c5.txt 2>&1; head -24 gpurun_out/bench_c5.txt
