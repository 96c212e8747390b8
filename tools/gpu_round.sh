#!/bin/bash
# One GPU session: parity tests, smoke, A/B, bench, ncu launch list + full capture of the raster kernels.
mkdir -p gpurun_out
nvidia-smi -L; nproc; lscpu | grep "Model name" | head -1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python tools/ab_forward.py c3 2>&1 | tee gpurun_out/ab_c3.json | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload c2 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cut -c1-400 gpurun_out/bench_c2.json
timeout 600 python bench.py --steps 5 --warmup 3 --workload c4 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; cat gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
timeout 300 python examples/demo2_deform.py --iters 200 2>&1 | tail -4 | tee gpurun_out/demo2.log
if [ "$1" != "noncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_softras -s 3 -c 3 -f -o gpurun_out/prof_softras python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
fi
ls gpurun_out
