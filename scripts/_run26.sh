mkdir -p gpurun_out/r01g
T="timeout 900 python -m pytest -q -m gpu --tb=short -p no:cacheprovider -x"
$T tests/test_gpu_kernels.py -k "gemm or stem" > gpurun_out/t_kern.log 2>&1; echo "kernel tests rc=$?"; tail -n 4 gpurun_out/t_kern.log
for d in 0 1; do echo "== BTSB_GEMM_DBG=$d"; BTSB_GEMM_DBG=$d timeout 300 python scripts/kbench.py --only "gemm" 2>&1 | grep gemm; done | tee gpurun_out/kbench_gemm_dbg.log
cap() { # name regex skip count cmd...
  n=$1; rx=$2; s=$3; c=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $s -c $c -o gpurun_out/r01g/$n -f "$@" > gpurun_out/r01g/ncu_$n.log 2>&1; echo "ncu $n rc=$?"
  python scripts/ncu_summary.py gpurun_out/r01g/$n.ncu-rep > gpurun_out/r01g/$n.summary.txt 2>&1
  ncu -i gpurun_out/r01g/$n.ncu-rep --page source --csv > gpurun_out/r01g/$n.source.csv 2>/dev/null
  gzip -f gpurun_out/r01g/$n.source.csv
}
cap fc1_bias gemm_tc_kernel 5 1 python scripts/kbench.py --only "fc1+bias C=320"
cap fc2_320 gemm_tc_kernel 5 1 python scripts/kbench.py --only "fc2+res C=320"
cat gpurun_out/r01g/*.summary.txt
find gpurun_out/r01g -name '*.ncu-rep' -size +6M -delete
