mkdir -p gpurun_out
for m in 0 1; do
  echo "== BTSB_GEMM_2CTA=$m"
  BTSB_GEMM_2CTA=$m timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_p$m.log 2>&1
  python scripts/show_bench.py gpurun_out/bench_c3_p$m.log 2>/dev/null | head -1 | cut -c1-120
  BTSB_GEMM_2CTA=$m timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_p$m.log 2>&1
  python scripts/show_bench.py gpurun_out/bench_c4_p$m.log 2>/dev/null | head -1 | cut -c1-120
done
