mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_training.py -q -m gpu -p no:cacheprovider -rA --tb=short > gpurun_out/t_train.log 2>&1; echo "pytest training rc=$?"
grep -E "parity|passed|failed|Error|error" gpurun_out/t_train.log | tail -n 25
timeout 600 python bench.py --workload c5 --steps 10 --warmup 3 > gpurun_out/bench_c5.log 2>&1; echo "bench c5 rc=$?"
python scripts/show_bench.py gpurun_out/bench_c5.log > gpurun_out/bench_c5.txt 2>&1; head -30 gpurun_out/bench_c5.txt
timeout 600 python bench.py --workload c5 --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_fp32.log 2>&1; echo "bench c5 fp32 rc=$?"
python scripts/show_bench.py gpurun_out/bench_c5_fp32.log > gpurun_out/bench_c5_fp32.txt 2>&1; head -12 gpurun_out/bench_c5_fp32.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3.log 2>&1; echo "bench c3 rc=$?"
python scripts/show_bench.py gpurun_out/bench_c3.log > gpurun_out/bench_c3.txt 2>&1; head -8 gpurun_out/bench_c3.txt
tail -c 600 gpurun_out/bench_c5.log
