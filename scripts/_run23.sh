mkdir -p gpurun_out/r01f
T="timeout 900 python -m pytest -q -m gpu --tb=short -p no:cacheprovider -x"
$T tests/test_gpu_kernels.py > gpurun_out/t_kern.log 2>&1; echo "kernel tests rc=$?"; tail -n 8 gpurun_out/t_kern.log
$T tests/test_gpu_models.py > gpurun_out/t_models.log 2>&1; echo "model tests rc=$?"; tail -n 5 gpurun_out/t_models.log
timeout 300 python scripts/kbench.py > gpurun_out/kbench.log 2>&1; echo "kbench rc=$?"; cat gpurun_out/kbench.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.log 2>&1; echo "bench c3 rc=$?"
python scripts/show_bench.py gpurun_out/bench_c3.log | head -24
cap() { # name regex skip count cmd...
  n=$1; rx=$2; s=$3; c=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $s -c $c -o gpurun_out/r01f/$n -f "$@" > gpurun_out/r01f/ncu_$n.log 2>&1; echo "ncu $n rc=$?"
  python scripts/ncu_summary.py gpurun_out/r01f/$n.ncu-rep > gpurun_out/r01f/$n.summary.txt 2>&1
  ncu -i gpurun_out/r01f/$n.ncu-rep --page source --csv > gpurun_out/r01f/$n.source.csv 2>/dev/null
  gzip -f gpurun_out/r01f/$n.source.csv
}
cap mlp2_80 mlp_fused2_kernel 5 1 python scripts/kbench.py --only "mlp_fused C=80"
cap fc1_320 gemm_tc_kernel 5 1 python scripts/kbench.py --only "fc1+gelu C=320"
cap fc2_320 gemm_tc_kernel 5 1 python scripts/kbench.py --only "fc2+res C=320"
cap dwln15 dwln3_kernel 5 1 python scripts/kbench.py --only "dwln 15x15"
cat gpurun_out/r01f/*.summary.txt
find gpurun_out/r01f -name '*.ncu-rep' -size +6M -delete
du -sh gpurun_out
