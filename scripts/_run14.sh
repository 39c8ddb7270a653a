mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu --tb=short -p no:cacheprovider -rA tests/test_gpu_maxvit.py > gpurun_out/t_maxvit.log 2>&1; echo "maxvit tests rc=$?"
grep -E "parity|passed|failed|FAILED|Error|error" gpurun_out/t_maxvit.log | head -60
