#!/bin/bash
# Optimistic (branch-free) divisions in forward/backward + fused mesh losses: parity suite, A/B against the guarded
# build, C3 / C5 / C2 bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
AB_WORKLOAD=c3 bash tools/ab_variants.sh 2>&1 | tee gpurun_out/ab_variants_c3.log | grep -v Warning
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cut -c1-330 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
timeout 600 python bench.py --workload c5 --steps 60 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c5.json')); print(d['modes'])"
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cut -c1-330 gpurun_out/bench_c2.json
