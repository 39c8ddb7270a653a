#!/usr/bin/env bash
# First visit of the next round (gpurun -- 'bash scripts/gpu_next.sh <tag>'): the default configuration's own bench line, and
# the unmeasured opt-in BTSB_MLP_V8=1 (256-bit residual / output accesses in the wide fused-MLP D2 epilogue) -- parity of the
# kernels and models under the switch, C3 bench A/B and the C = 320 pipeline trace under the switch.
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python bench.py --steps 20 --warmup 3 > $OUT/bench_c3.log 2>&1; echo "bench c3 rc=$?"
BTSB_MLP_V8=1 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -x -q -m gpu -p no:cacheprovider -rA > $OUT/t_v8.log 2>&1; echo "pytest v8 rc=$?"; tail -n 2 $OUT/t_v8.log
BTSB_MLP_V8=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_c3_v8.log 2>&1; echo "bench v8 rc=$?"
BTSB_MLP_V8=1 timeout 90 python scripts/mlp_trace.py 320 9 > $OUT/mlp_trace_320_v8.txt 2>&1; echo "trace rc=$?"
for f in bench_c3 bench_c3_v8; do python scripts/show_bench.py $OUT/$f.log 2>/dev/null | sed -n 1,6p | cut -c1-150; done
