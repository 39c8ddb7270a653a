#!/usr/bin/env bash
# Round-2 visit n: dw7x7+LN with two output rows per conv thread (BTSB_DWLN5_R2=1): parity + A/B; ingest on the GPU box.
OUT=gpurun_out/r02n
mkdir -p $OUT
BTSB_DWLN5_R2=1 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -q -m gpu -p no:cacheprovider -k "dwln or logits or intermediates" > $OUT/t_r2.log 2>&1; echo "pytest r2 rc=$?"; tail -n 3 $OUT/t_r2.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_c3.log 2>$OUT/bench_c3.err; echo "bench rc=$?"
BTSB_DWLN5_R2=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_c3_r2.log 2>$OUT/bench_c3_r2.err; echo "bench r2 rc=$?"
for f in bench_c3 bench_c3_r2; do python scripts/show_bench.py $OUT/$f.log 2>/dev/null | cut -c1-170 | sed -n 1,8p; done
timeout 300 python -m pytest tests/test_gpu_preprocess.py tests/test_host_api.py -q -p no:cacheprovider > $OUT/t_pre.log 2>&1; echo "pytest preprocess+host rc=$?"; tail -n 2 $OUT/t_pre.log
