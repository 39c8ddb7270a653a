#!/usr/bin/env bash
# visit: cubic-tanh GELU build (-DBTSB_GELU_CUBIC) against the default quintic: parity + kernel times.
OUT=gpurun_out/r02cu
mkdir -p $OUT
for v in cubic base; do
  if [ $v = base ]; then lib=""; else lib="BTSB_LIB=$PWD/btsbot_b200/libbtsbot_b200_cubic.so"; fi
  env $lib timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py tests/test_gpu_maxvit.py -q -m gpu -p no:cacheprovider -rA -k "mlp_fused_tcgen05 or bf16_logits or intermediates or fused_and" > $OUT/t_$v.log 2>&1; echo "$v pytest rc=$?"; tail -n 1 $OUT/t_$v.log; grep -E "^(FAILED|ERROR)" $OUT/t_$v.log | head -5
  grep "\[parity\].*bf16: gain\|fused vs\|intermediates" $OUT/t_$v.log | cut -c1-120
  env $lib BTSB_HOST_PACK=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c3_$v.log 2>$OUT/bench_c3_$v.err; echo "bench $v rc=$?"
  python scripts/show_bench.py $OUT/bench_c3_$v.log 2>/dev/null | cut -c1-150 | grep -E "value|mlp_fused"
done
