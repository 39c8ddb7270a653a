#!/usr/bin/env bash
# Round-2 visit j (2 GPUs): attention A/B on one GPU first (cheap), then the multi-rank bench lines.
OUT=gpurun_out/r02j
mkdir -p $OUT
CUDA_VISIBLE_DEVICES=0 timeout 200 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_c4.log 2>$OUT/bench_c4.err; echo "bench c4 rc=$?"
python scripts/show_bench.py $OUT/bench_c4.log 2>/dev/null | grep -E "value|attn" | cut -c1-170
CUDA_VISIBLE_DEVICES=0 timeout 200 python -m pytest tests/test_gpu_maxvit.py tests/test_gpu_models.py -q -m gpu -p no:cacheprovider -k "attention or bf16_logits" > $OUT/t_attn.log 2>&1; echo "pytest attn rc=$?"; tail -n 2 $OUT/t_attn.log
bash scripts/gpu_multi.sh 2 r02j_n2
