#!/usr/bin/env bash
# Round-2 visit b: defaults after the r02a A/B (EP=2 slabs, cp.async taps), and the y-buffer staging variant (EP=3).
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 60 scripts/micro/micro_fma2 > $OUT/micro_fma2.txt 2>&1; echo "micro_fma2 rc=$?"; cat $OUT/micro_fma2.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -x -q -m gpu -p no:cacheprovider > $OUT/t_default.log 2>&1; echo "pytest default rc=$?"; tail -n 2 $OUT/t_default.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_c3.log 2>&1; echo "bench rc=$?"
BTSB_MLP_EP=3 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -x -q -m gpu -p no:cacheprovider > $OUT/t_ep3.log 2>&1; echo "pytest ep3 rc=$?"; tail -n 5 $OUT/t_ep3.log
BTSB_MLP_EP=3 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_c3_ep3.log 2>&1; echo "bench ep3 rc=$?"
BTSB_MLP_EP=3 timeout 90 python scripts/mlp_trace.py 320 9 > $OUT/mlp_trace_320_ep3.txt 2>&1; echo "trace ep3 rc=$?"
for f in bench_c3 bench_c3_ep3; do python scripts/show_bench.py $OUT/$f.log 2>/dev/null | sed -n 1,9p | cut -c1-150; done
