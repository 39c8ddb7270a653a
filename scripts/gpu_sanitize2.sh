#!/usr/bin/env bash
# compute-sanitizer memcheck over the kernel tests of the kernels changed in the second half of round 2 (fp16 residual
# stream, in-place wide MLP, dedicated D2 warps, dwln5 at the pico widths, bf16-input K1): gpurun -- 'bash scripts/gpu_sanitize2.sh'
OUT=gpurun_out/r02san2
mkdir -p $OUT
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 1300 $SAN --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck.log \
  python -m pytest tests/test_gpu_kernels.py tests/test_gpu_preprocess.py -x -q -m gpu -p no:cacheprovider \
  -k "mlp_fused or dwln or gemm_bf16 or stem or lnpatch or saturates or cast or host_pack" > $OUT/memcheck_pytest.log 2>&1
echo "memcheck rc=$?"; tail -n 3 $OUT/memcheck.log; tail -n 2 $OUT/memcheck_pytest.log
