#!/usr/bin/env bash
# Builds btsbot_b200/libbtsbot_b200.so for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libbtsbot_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
       --expt-relaxed-constexpr -Xptxas -v ${BTSB_NVCC_EXTRA:-})
mkdir -p "$HERE/build"
OBJS=()
pids=()
for f in "$HERE"/*.cu; do
  o="$HERE/build/$(basename "${f%.cu}").o"
  OBJS+=("$o")
  if [[ ! -f "$o" || "$f" -nt "$o" || -n "$(find "$HERE" -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$o" 2>/dev/null)" \
        || "$HERE/../../include/btsbot_b200.h" -nt "$o" ]]; then
    ( "$NVCC" "${FLAGS[@]}" -c "$f" -o "$o" > "$o.log" 2>&1 || { cat "$o.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" -shared -o "$OUT" "${OBJS[@]}" -lcudart -lz
echo "built $OUT"
