#!/usr/bin/env bash
# Two-GPU visit (gpurun --gpus 2 -- 'bash scripts/gpu_multi.sh'): DDP == large-batch check, sharded scoring and DDP training.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tests/ddp_check.py > gpurun_out/ddp_check.log 2>&1; echo "ddp_check rc=$?"; tail -n 2 gpurun_out/ddp_check.log
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_n2.log 2>&1; echo "bench c3 n2 rc=$?"
timeout 600 $TR bench.py --gpus 2 --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_n2.log 2>&1; echo "bench c5 n2 rc=$?"
for f in bench_c3_n2 bench_c5_n2; do python scripts/show_bench.py gpurun_out/$f.log 2>/dev/null | sed -n 1,2p | cut -c1-200; done
