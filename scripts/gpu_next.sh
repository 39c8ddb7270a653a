#!/usr/bin/env bash
# First visit of the next round (gpurun -- 'bash scripts/gpu_next.sh <tag>'): the default configuration's own bench line, and
# the two unmeasured opt-in D2-epilogue variants of the wide fused MLP (BTSB_MLP_EP=1: 256-bit per-thread accesses,
# BTSB_MLP_EP=2: per-warp slabs + bulk tensor copies) -- parity of the kernels and models under each switch, C3 bench A/B and
# the C = 320 pipeline trace.
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python bench.py --steps 20 --warmup 3 > $OUT/bench_c3.log 2>&1; echo "bench c3 rc=$?"
for ep in 1 2; do
  BTSB_MLP_EP=$ep timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -x -q -m gpu -p no:cacheprovider -rA > $OUT/t_ep$ep.log 2>&1; echo "pytest ep$ep rc=$?"; tail -n 2 $OUT/t_ep$ep.log
  BTSB_MLP_EP=$ep timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_c3_ep$ep.log 2>&1; echo "bench ep$ep rc=$?"
  BTSB_MLP_EP=$ep timeout 90 python scripts/mlp_trace.py 320 9 > $OUT/mlp_trace_320_ep$ep.txt 2>&1; echo "trace ep$ep rc=$?"
done
# cp.async tap staging in the 3x3 / 1x1 dw+LN kernel
BTSB_DWS_CPA=1 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -x -q -m gpu -p no:cacheprovider > $OUT/t_cpa.log 2>&1; echo "pytest cpa rc=$?"; tail -n 2 $OUT/t_cpa.log
BTSB_DWS_CPA=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_c3_cpa.log 2>&1; echo "bench cpa rc=$?"
for f in bench_c3 bench_c3_ep1 bench_c3_ep2 bench_c3_cpa; do python scripts/show_bench.py $OUT/$f.log 2>/dev/null | sed -n 1,9p | cut -c1-150; done
