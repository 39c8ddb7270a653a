mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/ddp_check.py > gpurun_out/ddp.log 2>&1; echo "ddp rc=$?"
grep -v "^$" gpurun_out/ddp.log | grep -E "Error|error|assert|Traceback|File \"/|DDP|world" | head -30
