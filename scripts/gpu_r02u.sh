#!/usr/bin/env bash
# Round-2 visit u: fp16 residual stream (BF16_XF16) + in-place wide MLP by default + quintic-sat GELU: parity + bench A/B.
OUT=gpurun_out/r02u
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -rA -x > $OUT/t_k.log 2>&1; echo "pytest kernels rc=$?"; tail -n 2 $OUT/t_k.log; grep -E "^(FAILED|ERROR)" $OUT/t_k.log | head
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_maxvit.py -q -m gpu -p no:cacheprovider -rA > $OUT/t_m.log 2>&1; echo "pytest models rc=$?"; tail -n 2 $OUT/t_m.log; grep -E "^(FAILED|ERROR)" $OUT/t_m.log | head; grep "\[parity\].*bf16: gain" $OUT/t_m.log | cut -c1-200; grep "\[parity\] fused vs\|intermediates\|large batch" $OUT/t_m.log | cut -c1-200
BTSB_XF16=0 BTSB_MLP_INPLACE=0 timeout 600 python -m pytest tests/test_gpu_models.py -q -m gpu -p no:cacheprovider -rA -k "bf16 or fused_and" > $OUT/t_m_bf16stream.log 2>&1; echo "pytest models (bf16 stream, out of place) rc=$?"; tail -n 2 $OUT/t_m_bf16stream.log; grep "\[parity\].*bf16: gain\|fused vs" $OUT/t_m_bf16stream.log | cut -c1-200
for v in default bf16stream outofplace; do
  case $v in
    default) envs="BTSB_X=0";;
    bf16stream) envs="BTSB_XF16=0";;
    outofplace) envs="BTSB_MLP_INPLACE=0";;
  esac
  env $envs timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c3_$v.log 2>$OUT/bench_c3_$v.err; echo "bench $v rc=$?"; tail -n 2 $OUT/bench_c3_$v.err
  python scripts/show_bench.py $OUT/bench_c3_$v.log 2>/dev/null | cut -c1-170 | sed -n 1,17p
done
timeout 90 python scripts/mlp_trace.py 320 9 > $OUT/mlp_trace_320.txt 2>&1
BTSB_MLP_INPLACE=1 timeout 90 python scripts/mlp_trace.py 320 9 > $OUT/mlp_trace_320_inplace.txt 2>&1
