mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
for wl in c3 c2; do
  echo "== $wl"; AB_NO_REF=1 timeout 300 python tools/ab_forward.py $wl 2>&1 | tail -1 | cut -c90-330
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cut -c1-250 gpurun_out/bench_c3.json
