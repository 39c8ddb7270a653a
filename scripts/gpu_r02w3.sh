#!/usr/bin/env bash
# visit: warp-per-image 3x3 dw+LN kernel: parity + A/B.
OUT=gpurun_out/r02w3
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -q -m gpu -p no:cacheprovider -k "test_dwln or bf16_logits or intermediates" > $OUT/t.log 2>&1; echo "pytest rc=$?"; tail -n 2 $OUT/t.log; grep -E "^(FAILED|ERROR)" $OUT/t.log | head
for v in w3 small; do
  if [ $v = w3 ]; then envs="BTSB_X=0"; else envs="BTSB_DWLN_W3=0"; fi
  env $envs BTSB_HOST_PACK=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c3_$v.log 2>$OUT/bench_c3_$v.err; echo "bench $v rc=$?"; tail -n 2 $OUT/bench_c3_$v.err
  python scripts/show_bench.py $OUT/bench_c3_$v.log 2>/dev/null | cut -c1-150 | grep -E "value|dwln_3|dwln_1"
done
