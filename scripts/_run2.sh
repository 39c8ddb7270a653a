T="timeout 900 python -m pytest -q -m gpu --tb=short -rA -p no:cacheprovider"
$T tests > gpurun_out/t_all.log 2>&1; echo "all rc=$?"
tail -n 4 gpurun_out/t_all.log
bash scripts/gpu_profile.sh r01a dwln_kernel gemm_tc_kernel stem_kernel
