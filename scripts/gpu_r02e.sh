#!/usr/bin/env bash
# Round-2 visit e: full suite (no -x), then the CTA-pair wide fused MLP (BTSB_MLP_PAIR=1): parity + A/B bench.
TAG=${1:-r02e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/ -q -m gpu -p no:cacheprovider -rA > $OUT/t_all.log 2>&1; echo "pytest gpu rc=$?"; tail -n 3 $OUT/t_all.log; grep -E "^(FAILED|ERROR)" $OUT/t_all.log | head
grep -h "^\[parity\]" $OUT/t_all.log > $OUT/parity_lines.txt
BTSB_MLP_PAIR=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "mlp_fused" -x > $OUT/t_pair_k.log 2>&1; echo "pytest pair kernels rc=$?"; tail -n 4 $OUT/t_pair_k.log
BTSB_MLP_PAIR=1 timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_maxvit.py -q -m gpu -p no:cacheprovider > $OUT/t_pair_m.log 2>&1; echo "pytest pair models rc=$?"; tail -n 3 $OUT/t_pair_m.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_c3.log 2>$OUT/bench_c3.err; echo "bench rc=$?"
BTSB_MLP_PAIR=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_c3_pair.log 2>$OUT/bench_c3_pair.err; echo "bench pair rc=$?"; tail -n 3 $OUT/bench_c3_pair.err
for f in bench_c3 bench_c3_pair; do python scripts/show_bench.py $OUT/$f.log 2>/dev/null | cut -c1-170 | sed -n 1,8p; done
