#!/usr/bin/env bash
# A/B visit: full GPU suite on the defaults, the kernels touched by the opt-in flags re-tested with the flags on, and the
# C3 bench under each flag setting (gpurun -- 'bash scripts/gpu_ab.sh').
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider -rA > gpurun_out/t_all.log 2>&1; echo "pytest gpu rc=$?"; tail -n 2 gpurun_out/t_all.log
grep -h "^\[parity\]" gpurun_out/t_all.log > gpurun_out/parity_lines.txt
BTSB_DWS_PF=2 BTSB_HEAD_TC=1 timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -rA > gpurun_out/t_flags.log 2>&1; echo "pytest flags rc=$?"; tail -n 2 gpurun_out/t_flags.log
grep -h "head feature" gpurun_out/t_all.log gpurun_out/t_flags.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_base.log 2>&1; echo "base rc=$?"
BTSB_DWS_PF=2 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_pf2.log 2>&1; echo "pf2 rc=$?"
BTSB_DWS_PF=2 BTSB_HEAD_TC=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_pf2_head.log 2>&1; echo "pf2+head rc=$?"
for w in base pf2 pf2_head; do echo "== $w"; python scripts/show_bench.py gpurun_out/ab_$w.log 2>/dev/null | sed -n 1,16p | cut -c1-150; done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 16384 > gpurun_out/ab_b16k.log 2>&1; echo "b16k rc=$?"
python scripts/show_bench.py gpurun_out/ab_b16k.log 2>/dev/null | sed -n 1,1p | cut -c1-150
timeout 90 python scripts/mlp_trace.py 320 9 > gpurun_out/mlp_trace_320.txt 2>&1; echo "trace rc=$?"
