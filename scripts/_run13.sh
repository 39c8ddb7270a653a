mkdir -p gpurun_out
T="timeout 900 python -m pytest -q -m gpu --tb=short -p no:cacheprovider -x"
$T tests > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_bf16.log 2>&1; echo "bench bf16 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"
tail -n 15 gpurun_out/t_all.log
tail -n 5 gpurun_out/smoke.log
tail -c 3000 gpurun_out/bench_bf16.log
tail -c 800 gpurun_out/bench_ref.log
