mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "tcgen05 or mlp or fused" -p no:cacheprovider --tb=short > gpurun_out/t_tc.log 2>&1; echo "tc tests rc=$?"; tail -n 3 gpurun_out/t_tc.log
timeout 300 python scripts/kbench.py --only mlp_fused 2>&1 | grep mlp_fused | tee gpurun_out/kbench_mlp.log
timeout 300 python scripts/mlp_trace.py 80 225 > gpurun_out/mlp_trace_80.txt 2>&1; echo "trace80 rc=$?"
timeout 300 python scripts/mlp_trace.py 160 49 > gpurun_out/mlp_trace_160.txt 2>&1; echo "trace160 rc=$?"
