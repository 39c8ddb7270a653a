mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tests/ddp_check.py > gpurun_out/ddp_check.log 2>&1; echo "ddp_check rc=$?"; tail -n 3 gpurun_out/ddp_check.log
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_n2.log 2>&1; echo "bench c3 n2 rc=$?"
python scripts/show_bench.py gpurun_out/bench_c3_n2.log | head -2 | cut -c1-300
timeout 600 $TR bench.py --gpus 2 --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_n2.log 2>&1; echo "bench c5 n2 rc=$?"
python scripts/show_bench.py gpurun_out/bench_c5_n2.log | head -2 | cut -c1-300
timeout 600 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_n1.log 2>&1; echo "bench c5 n1 rc=$?"
python scripts/show_bench.py gpurun_out/bench_c5_n1.log | head -1 | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_train_loop.py -q -m gpu -p no:cacheprovider -rs 2>&1 | tail -4
