#!/usr/bin/env bash
# Round-2 first visit: A/B of the unmeasured round-1 variants, then compute-sanitizer (bounded).
bash scripts/gpu_next.sh r02a
sed -i 's/timeout 1500/timeout 700/; s/timeout 900/timeout 500/' scripts/gpu_sanitize.sh
bash scripts/gpu_sanitize.sh r02a_san
