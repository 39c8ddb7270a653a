#!/usr/bin/env bash
# visit: fp16 -> fp32 conversion variants in the dw7x7+LN kernel (BTSB_XCVT 0 / 1 / 2): dwln parity + kernel times.
OUT=gpurun_out/r02cv
mkdir -p $OUT
for v in 0 1 2; do
  if [ $v = 0 ]; then lib=""; else lib="BTSB_LIB=$PWD/btsbot_b200/libbtsbot_b200_xcvt$v.so"; fi
  env $lib timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "test_dwln" > $OUT/t_dw_$v.log 2>&1; echo "xcvt $v pytest rc=$?"; tail -n 1 $OUT/t_dw_$v.log
  env $lib BTSB_HOST_PACK=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c3_$v.log 2>$OUT/bench_c3_$v.err; echo "bench $v rc=$?"
  python scripts/show_bench.py $OUT/bench_c3_$v.log 2>/dev/null | cut -c1-150 | grep -E "value|dwln_15|dwln_7"
done
