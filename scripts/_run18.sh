# focused source-level ncu captures (small reports), summarised on the box
mkdir -p gpurun_out/r01c
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
cap() { # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/r01c/$1 -f $BENCH > gpurun_out/r01c/ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
  python scripts/ncu_summary.py gpurun_out/r01c/$1.ncu-rep > gpurun_out/r01c/$1.summary.txt 2>&1
  ncu -i gpurun_out/r01c/$1.ncu-rep --page source --csv > gpurun_out/r01c/$1.source.csv 2>/dev/null
  ncu -i gpurun_out/r01c/$1.ncu-rep --page details > gpurun_out/r01c/$1.details.txt 2>/dev/null
  ls -la gpurun_out/r01c/$1.ncu-rep
}
cap mlp_fused mlp_fused_kernel 8 1
cap mlp_fused160 mlp_fused_kernel 10 1
cap gemm_fc gemm_tc_kernel 75 2
cap dwln3 dwln3_kernel 12 3
cap dwln2 dwln2_kernel 30 1
cap misc "lnpatch|meta_head" 8 4
du -sh gpurun_out/r01c
# keep the transfer under the 64 MiB cap
find gpurun_out/r01c -name '*.ncu-rep' -size +9M -delete
gzip -f gpurun_out/r01c/*.source.csv
du -sh gpurun_out
