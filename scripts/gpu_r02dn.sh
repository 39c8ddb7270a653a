#!/usr/bin/env bash
# visit: fused downsample kernel: kernel tests, model parity, bench A/B.
OUT=gpurun_out/r02dn
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -rA -x -k "down_fused or saturates" > $OUT/t_k.log 2>&1; echo "pytest kernels rc=$?"; tail -n 2 $OUT/t_k.log; grep -E "^(FAILED|ERROR)" $OUT/t_k.log | head; grep "down_fused" $OUT/t_k.log | grep parity | head -12 | cut -c1-160
timeout 900 python -m pytest tests/test_gpu_models.py -q -m gpu -p no:cacheprovider -rA > $OUT/t_m.log 2>&1; echo "pytest models rc=$?"; tail -n 2 $OUT/t_m.log; grep -E "^(FAILED|ERROR)" $OUT/t_m.log | head; grep "\[parity\].*bf16: gain" $OUT/t_m.log | cut -c1-200
for v in fused unfused; do
  if [ $v = fused ]; then envs="BTSB_X=0"; else envs="BTSB_DOWN_FUSED=0"; fi
  env $envs BTSB_HOST_PACK=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c3_$v.log 2>$OUT/bench_c3_$v.err; echo "bench $v rc=$?"; tail -n 2 $OUT/bench_c3_$v.err
  python scripts/show_bench.py $OUT/bench_c3_$v.log 2>/dev/null | cut -c1-170 | sed -n 1,16p
done
