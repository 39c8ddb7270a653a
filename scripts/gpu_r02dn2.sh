#!/usr/bin/env bash
OUT=gpurun_out/r02dn
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --alerts 0"
BTSB_HOST_PACK=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:down_fused_kernel -s 2 -c 1 -o $OUT/down_fused -f $BENCH > $OUT/ncu_down.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py $OUT/down_fused.ncu-rep > $OUT/down_fused.summary.txt 2>/dev/null; cat $OUT/down_fused.summary.txt | cut -c1-600
