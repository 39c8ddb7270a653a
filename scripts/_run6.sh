mkdir -p gpurun_out
nvidia-smi -L
T="timeout 900 python -m pytest -q -m gpu --tb=short -rA -p no:cacheprovider"
$T tests/test_gpu_train_loop.py tests/test_gpu_training.py > gpurun_out/t_train.log 2>&1; echo "train rc=$?"; grep -E "^\[|parity|passed|failed|Error|assert|DDP|world" gpurun_out/t_train.log | head -40
for N in 1 2; do
  if [ $N = 1 ]; then timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.log 2>&1;
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.log 2>&1; fi
  echo "bench N=$N rc=$?"; python scripts/show_bench.py gpurun_out/bench_n$N.log 2>/dev/null | head -2
done
