#!/bin/bash
# One GPU session: parity tests, smoke, bench, ncu launch list + full capture of the raster kernels.
mkdir -p gpurun_out
nvidia-smi -L; nproc; lscpu | grep "Model name" | head -1
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
python bench.py --steps 10 --warmup 3 --workload c2 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
if [ "$1" != "noncu" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_softras -s 2 -c 2 -f -o gpurun_out/prof_softras python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
fi
if [ -f oracle/_ref/libjrender_ref.so ]; then python -m oracle.make_ref_golden > gpurun_out/golden.log 2>&1; tail -12 gpurun_out/golden.log; fi
ls gpurun_out
