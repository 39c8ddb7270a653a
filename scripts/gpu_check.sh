#!/usr/bin/env bash
# One GPU-box visit (gpurun -- 'bash scripts/gpu_check.sh'): every GPU parity test, the smoke entry, and the three benches.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider -rA > gpurun_out/t_all.log 2>&1; echo "pytest gpu rc=$?"; tail -n 3 gpurun_out/t_all.log
grep -h "^\[parity\]" gpurun_out/t_all.log > gpurun_out/parity_lines.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "reference arm rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3.log 2>&1; echo "bench c3 rc=$?"
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.log 2>&1; echo "bench c4 rc=$?"
timeout 600 python bench.py --workload c5 --steps 10 --warmup 3 > gpurun_out/bench_c5.log 2>&1; echo "bench c5 rc=$?"
for w in c3 c4 c5; do python scripts/show_bench.py gpurun_out/bench_$w.log > gpurun_out/bench_$w.txt 2>/dev/null; sed -n 1,12p gpurun_out/bench_$w.txt | cut -c1-170; done
