#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
AB_WORKLOAD=c3 bash tools/ab_variants.sh 2>&1 | tee gpurun_out/ab_variants_c3.log | grep -v Warning
for lib in jrender_b200/lib/libb200raster*.so; do echo "== c5 $lib"; B200R_LIB=$PWD/$lib timeout 300 python tools/ab_c5.py 120 2>&1 | tail -1; done | tee gpurun_out/ab_variants_c5.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cut -c1-330 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
timeout 600 python bench.py --workload c5 --steps 60 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c5.json')); print(d['modes'])"
