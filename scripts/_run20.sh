mkdir -p gpurun_out/r01d
./scripts/micro/micro_gelu > gpurun_out/micro_gelu.log 2>&1; cat gpurun_out/micro_gelu.log
T="timeout 900 python -m pytest -q -m gpu --tb=short -p no:cacheprovider -x"
$T tests/test_gpu_kernels.py -k "gemm" > gpurun_out/t_kern.log 2>&1; echo "gemm tests rc=$?"; tail -n 3 gpurun_out/t_kern.log
cap() { # name regex skip count cmd...
  n=$1; rx=$2; s=$3; c=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $s -c $c -o gpurun_out/r01d/$n -f "$@" > gpurun_out/r01d/ncu_$n.log 2>&1; echo "ncu $n rc=$?"
  python scripts/ncu_summary.py gpurun_out/r01d/$n.ncu-rep > gpurun_out/r01d/$n.summary.txt 2>&1
  ncu -i gpurun_out/r01d/$n.ncu-rep --page source --csv > gpurun_out/r01d/$n.source.csv 2>/dev/null
  gzip -f gpurun_out/r01d/$n.source.csv
}
cap mlp2_80 mlp_fused2_kernel 5 1 python scripts/kbench.py --only "mlp_fused C=80"
cap mlp2_160 mlp_fused2_kernel 5 1 python scripts/kbench.py --only "mlp_fused C=160"
cap fc1_320 gemm_tc_kernel 5 1 python scripts/kbench.py --only "fc1+gelu C=320"
cap head meta_head 3 1 python scripts/kbench.py --only "meta_head"
cat gpurun_out/r01d/*.summary.txt
find gpurun_out/r01d -name '*.ncu-rep' -size +6M -delete
du -sh gpurun_out
