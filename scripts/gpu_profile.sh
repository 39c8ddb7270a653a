#!/usr/bin/env bash
# ncu evidence for one bench step: (1) launch list with per-launch device time, (2) --set full on the top kernels.
# Usage: scripts/gpu_profile.sh <tag> [kernel-regex ...]
TAG=${1:-r01}; shift
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --batch 4096 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
for K in "$@"; do
  N=$(echo "$K" | tr -c 'a-zA-Z0-9' '_')
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 2 \
      -o gpurun_out/prof_${TAG}_${N} -f $BENCH > gpurun_out/ncu_full_${TAG}_${N}.log 2>&1
  echo "full $K rc=$?"
done
ls -la gpurun_out | tail -20
