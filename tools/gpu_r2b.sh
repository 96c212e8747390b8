#!/bin/bash
# Round-2 session B: full GPU test-suite, NMR K9 occupancy A/B, forward round-size A/B.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
for v in "" _k9m4 _k9m6; do
  echo "== c4 lib$v"
  B200R_LIB=$PWD/jrender_b200/lib/libb200raster$v.so timeout 600 python bench.py --steps 5 --warmup 3 --workload c4 --no-reference-gpu 2> gpurun_out/bench_c4$v.err | tee gpurun_out/bench_c4$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()})"
done
AB_WORKLOAD=c3 bash tools/ab_variants.sh 2>&1 | grep -v Warning | tee gpurun_out/ab_variants_c3.log
