#!/usr/bin/env bash
# Last visit of the round: full GPU suite, smoke, reference arm, C3 bench (+ A/B of the early residual fetch and a
# 16384-alert step), and the ncu launch list of the default bench command.
TAG=${1:-r01n}
mkdir -p gpurun_out/$TAG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider -rA > gpurun_out/t_all.log 2>&1; echo "pytest gpu rc=$?"; tail -n 2 gpurun_out/t_all.log
grep -h "^\[parity\]" gpurun_out/t_all.log > gpurun_out/parity_lines.txt
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "reference arm rc=$?"
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3.log 2>&1; echo "bench c3 rc=$?"
BTSB_MLP_RES_EARLY=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_res_late.log 2>&1; echo "res late rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 16384 > gpurun_out/ab_b16k.log 2>&1; echo "b16k rc=$?"
python scripts/show_bench.py gpurun_out/bench_c3.log > gpurun_out/bench_c3.txt 2>/dev/null; sed -n 1,18p gpurun_out/bench_c3.txt | cut -c1-150
for w in res_late b16k; do python scripts/show_bench.py gpurun_out/ab_$w.log 2>/dev/null | sed -n 1,3p | cut -c1-150; done
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/$TAG/launches_c3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/$TAG/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 60 python scripts/mlp_trace.py 320 9 > gpurun_out/$TAG/mlp_trace_320_early.txt 2>&1; echo "trace rc=$?"
