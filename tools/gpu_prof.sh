#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_softras_gpu.py -q -m gpu -x 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_softras_forward -s 1 -c 1 -f -o gpurun_out/prof_c3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1; tail -1 gpurun_out/ncu_c3.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cut -c1-330 gpurun_out/bench_c3.json
