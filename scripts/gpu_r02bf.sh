#!/usr/bin/env bash
# visit: fc1 bias folded into G1 of the narrow fused MLP (C = 64 / 80): parity + A/B.
OUT=gpurun_out/r02bf
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -rA -x -k "mlp_fused" > $OUT/t_k.log 2>&1; echo "pytest kernels rc=$?"; tail -n 2 $OUT/t_k.log; grep -E "^(FAILED|ERROR)" $OUT/t_k.log | head; grep "bias-folded" $OUT/t_k.log | grep parity | cut -c1-150
timeout 900 python -m pytest tests/test_gpu_models.py -q -m gpu -p no:cacheprovider -rA > $OUT/t_m.log 2>&1; echo "pytest models rc=$?"; tail -n 2 $OUT/t_m.log; grep -E "^(FAILED|ERROR)" $OUT/t_m.log | head; grep "\[parity\].*bf16: gain" $OUT/t_m.log | cut -c1-130
for v in folded unfolded; do
  if [ $v = folded ]; then envs="BTSB_X=0"; else envs="BTSB_FOLD_BIAS=0"; fi
  env $envs BTSB_HOST_PACK=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c3_$v.log 2>$OUT/bench_c3_$v.err; echo "bench $v rc=$?"; tail -n 2 $OUT/bench_c3_$v.err
  python scripts/show_bench.py $OUT/bench_c3_$v.log 2>/dev/null | cut -c1-150 | grep -E "value|mlp_fused"
done
