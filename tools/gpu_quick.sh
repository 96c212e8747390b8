#!/bin/bash
# Short GPU session: parity tests, forward A/B, C3 bench, optional extra command ($1).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python tools/ab_forward.py c3 2>&1 | tee gpurun_out/ab_c3.json | tail -3 | cut -c1-1500
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cut -c1-700 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
if [ -n "$1" ]; then bash -c "$1"; fi
