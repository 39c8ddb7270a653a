#!/usr/bin/env bash
# Round-2 visit t: cubic 2xGELU in the fused MLP (A/B against the quintic build), in-place wide MLP (EP=4), AlertScorer ring.
OUT=gpurun_out/r02t
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "mlp_fused" -rA > $OUT/t_k.log 2>&1; echo "pytest kernels rc=$?"; tail -n 2 $OUT/t_k.log; grep "in place" $OUT/t_k.log | head
timeout 900 python -m pytest tests/test_gpu_models.py -q -m gpu -p no:cacheprovider -rA > $OUT/t_m.log 2>&1; echo "pytest models rc=$?"; tail -n 2 $OUT/t_m.log; grep -E "^(FAILED|ERROR)" $OUT/t_m.log | head; grep "\[parity\].*bf16" $OUT/t_m.log | cut -c1-200
BTSB_MLP_INPLACE=1 timeout 600 python -m pytest tests/test_gpu_models.py -q -m gpu -p no:cacheprovider -rA -k "bf16 or intermediates" > $OUT/t_m_inplace.log 2>&1; echo "pytest models (in place) rc=$?"; tail -n 2 $OUT/t_m_inplace.log; grep "\[parity\].*bf16" $OUT/t_m_inplace.log | cut -c1-200
for v in default quintic inplace; do
  case $v in
    default) envs="";;
    quintic) envs="BTSB_LIB=$PWD/btsbot_b200/libbtsbot_b200_quintic.so";;
    inplace) envs="BTSB_MLP_INPLACE=1";;
  esac
  env $envs timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c3_$v.log 2>$OUT/bench_c3_$v.err; echo "bench $v rc=$?"; tail -n 2 $OUT/bench_c3_$v.err
  python scripts/show_bench.py $OUT/bench_c3_$v.log 2>/dev/null | cut -c1-170 | sed -n 1,7p
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c3_default2.log 2>$OUT/bench_c3_default2.err; python scripts/show_bench.py $OUT/bench_c3_default2.log 2>/dev/null | cut -c1-170 | sed -n 1,2p
BTSB_MLP_INPLACE=1 timeout 90 python scripts/mlp_trace.py 320 9 > $OUT/mlp_trace_320_inplace.txt 2>&1
timeout 90 python scripts/mlp_trace.py 320 9 > $OUT/mlp_trace_320.txt 2>&1
timeout 90 python scripts/mlp_trace.py 80 9 > $OUT/mlp_trace_80.txt 2>&1
