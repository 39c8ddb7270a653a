#!/usr/bin/env bash
# Round-2 visit f: the 3xTF32 tensor-core GEMM of the fp32 mode -- parity (fp32 kernel + model tests), fp32 C3 bench A/B.
TAG=${1:-r02f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "gemm_f32 or lnpatch" -rA > $OUT/t_k.log 2>&1; echo "pytest kernels rc=$?"; tail -n 3 $OUT/t_k.log; grep "\[parity\].*gemm" $OUT/t_k.log | head -12
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_maxvit.py -q -m gpu -p no:cacheprovider -rA > $OUT/t_m.log 2>&1; echo "pytest models rc=$?"; tail -n 3 $OUT/t_m.log; grep -E "^(FAILED|ERROR)" $OUT/t_m.log | head; grep "\[parity\].*fp32" $OUT/t_m.log | head -20
timeout 300 python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_fp32.log 2>$OUT/bench_fp32.err; echo "bench fp32 rc=$?"; tail -n 3 $OUT/bench_fp32.err
BTSB_F32_SIMT=1 timeout 300 python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_fp32_simt.log 2>$OUT/bench_fp32_simt.err; echo "bench fp32 simt rc=$?"
for f in bench_fp32 bench_fp32_simt; do python scripts/show_bench.py $OUT/$f.log 2>/dev/null | cut -c1-170 | sed -n 1,22p; done
