#!/bin/bash
# end-of-round state: parity suite, smoke, every bench workload, ncu launch list + full capture (with source)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cut -c1-300 gpurun_out/bench_c3.json; tail -2 gpurun_out/bench_c3.err
timeout 600 python bench.py --workload c5 --steps 60 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c5.json')); print(d['value'], d['modes'])"
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cut -c1-200 gpurun_out/bench_c2.json
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; cut -c1-200 gpurun_out/bench_c4.json
timeout 300 python examples/demo2_deform.py --iters 200 --cuda-graph 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_softras -s 3 -c 3 -f -o gpurun_out/prof_softras python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1; tail -1 gpurun_out/ncu_c3.log
