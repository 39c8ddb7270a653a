mkdir -p gpurun_out
timeout 300 python scripts/mlp_trace.py 80 225 > gpurun_out/mlp_trace_80.txt 2>&1; echo "trace80 rc=$?"
timeout 300 python scripts/mlp_trace.py 160 49 > gpurun_out/mlp_trace_160.txt 2>&1; echo "trace160 rc=$?"
