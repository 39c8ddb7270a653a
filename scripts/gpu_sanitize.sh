#!/usr/bin/env bash
# compute-sanitizer pass over the kernel-level parity tests (gpurun -- 'bash scripts/gpu_sanitize.sh <tag>').
# memcheck on every kernel test (small shapes: minutes under the sanitizer), racecheck on the shared-memory heavy kernels.
# tcgen05 / TMA kernels: memcheck covers their global accesses; racecheck does not model the async proxy, so it is limited to
# the SIMT kernels (dw+LN, LayerNorm patches, head, preprocessing).
TAG=${1:-sanitize}
OUT=gpurun_out/$TAG
mkdir -p $OUT
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $SAN --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck.log \
  python -m pytest tests/test_gpu_kernels.py tests/test_gpu_preprocess.py -x -q -m gpu -p no:cacheprovider > $OUT/memcheck_pytest.log 2>&1
echo "memcheck rc=$?"; tail -n 3 $OUT/memcheck.log
timeout 900 $SAN --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck.log \
  python -m pytest tests/test_gpu_kernels.py tests/test_gpu_preprocess.py -x -q -m gpu -p no:cacheprovider -k "dwln or lnpatch or head or crop or pad" > $OUT/racecheck_pytest.log 2>&1
echo "racecheck rc=$?"; tail -n 3 $OUT/racecheck.log
