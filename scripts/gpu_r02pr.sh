#!/usr/bin/env bash
# visit: plain-vs-split probe of the e2e path + final conv-kernel conversion: tests + default-style bench lines.
OUT=gpurun_out/r02pr
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "alert_scorer or host_pack or dwln or bf16_logits" > $OUT/t.log 2>&1; echo "pytest rc=$?"; tail -n 2 $OUT/t.log; grep -E "^(FAILED|ERROR)" $OUT/t.log | head
for v in auto plain; do
  if [ $v = auto ]; then envs="BTSB_X=0"; else envs="BTSB_HOST_PACK=0"; fi
  env $envs timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_c3_$v.log 2>$OUT/bench_c3_$v.err; echo "bench $v rc=$?"; tail -n 2 $OUT/bench_c3_$v.err
  python - $OUT/bench_c3_$v.log <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); e = d["e2e"]
        print("   value %.0f  e2e %.0f  ms/step %.3f  h2d %.1f GB/s  host_pack %s" % (d["value"], e["value"], e["ms_per_step"], e["h2d_gbs_per_gpu"], e.get("host_pack")))
PY
done
