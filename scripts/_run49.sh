mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest gpu rc=$?"; tail -n 4 gpurun_out/t_all.log
for w in 0 1; do
  echo "== BTSB_FUSE_WIDE=$w"
  BTSB_FUSE_WIDE=$w timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_w$w.log 2>&1
  python scripts/show_bench.py gpurun_out/bench_c3_w$w.log > gpurun_out/bench_c3_w$w.txt 2>/dev/null; head -9 gpurun_out/bench_c3_w$w.txt | cut -c1-130
done
