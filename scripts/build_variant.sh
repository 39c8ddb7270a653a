#!/usr/bin/env bash
# Builds an A/B variant of the library: scripts/build_variant.sh <suffix> <file.cu> <extra nvcc flags...>
# -> btsbot_b200/libbtsbot_b200_<suffix>.so = the current objects with <file.cu> recompiled with the extra flags.
# Select it at run time with BTSB_LIB=btsbot_b200/libbtsbot_b200_<suffix>.so.
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
SUF="$1"; SRC="$2"; shift 2
CS="$ROOT/btsbot_b200/csrc"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
bash "$CS/build.sh" > /dev/null
O="$CS/build/${SRC%.cu}_$SUF.o"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
  -c "$CS/$SRC" -o "$O"
OBJS=()
for f in "$CS"/*.cu; do
  b="$(basename "${f%.cu}")"
  if [[ "$b.cu" == "$SRC" ]]; then OBJS+=("$O"); else OBJS+=("$CS/build/$b.o"); fi
done
"$NVCC" -shared -o "$ROOT/btsbot_b200/libbtsbot_b200_$SUF.so" "${OBJS[@]}" -lcudart -lz
echo "built btsbot_b200/libbtsbot_b200_$SUF.so"
