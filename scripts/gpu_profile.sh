#!/usr/bin/env bash
# ncu evidence for one C3 bench step (gpurun -- 'bash scripts/gpu_profile.sh <tag>'):
#   (1) launch list with per-launch device time (cold-cache, serialised: compare SHARES with bench.py's `kernels`),
#   (2) --set full + source on the top kernels.  Per step the fused-MLP launches are 2 x <80>, 2 x <160>, 8 x <320> and
#       the dw7x7+LN (dwln5) launches 2 x <15,80>, 2 x <7,160>; -s skips the warm-up steps' launches.
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --alerts 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_c3.csv $BENCH > $OUT/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mlp_fused2_kernel -s 28 -c 1 -o $OUT/mlp2_320 -f $BENCH > $OUT/ncu_mlp320.log 2>&1; echo "ncu mlp320 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mlp_fused2_kernel -s 24 -c 1 -o $OUT/mlp2_80 -f $BENCH > $OUT/ncu_mlp80.log 2>&1; echo "ncu mlp80 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dwln5_kernel -s 8 -c 1 -o $OUT/dwln15 -f $BENCH > $OUT/ncu_dw.log 2>&1; echo "ncu dwln rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cast_transpose63 -s 2 -c 1 -o $OUT/k1_cast -f $BENCH > $OUT/ncu_k1.log 2>&1; echo "ncu k1 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 20 -c 1 -o $OUT/tf32_gemm -f python bench.py --precision fp32 --steps 1 --warmup 3 --no-cpu-baseline --no-extras --alerts 0 > $OUT/ncu_tf32.log 2>&1; echo "ncu tf32 rc=$?"
for f in mlp2_320 mlp2_80 dwln15 k1_cast tf32_gemm; do python scripts/ncu_summary.py $OUT/$f.ncu-rep > $OUT/$f.summary.txt 2>/dev/null; cat $OUT/$f.summary.txt | cut -c1-400; done
python scripts/launch_shares.py $OUT/launches_c3.csv > $OUT/launch_shares.txt 2>/dev/null; head -n 25 $OUT/launch_shares.txt
ls -la $OUT | head -n 30
