#!/bin/bash
# Round-2 session A: new full-size parity tests, bench line with the new baseline legs, NMR backward ncu capture.
mkdir -p gpurun_out
nvidia-smi -L | head -2; nproc; lscpu | grep "Model name" | head -1
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cut -c1-600 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
timeout 900 python bench.py --steps 5 --warmup 3 --workload c4 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; cut -c1-900 gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nmr_backward_pixel_map -s 1 -c 1 -f -o gpurun_out/prof_nmr_bwd python bench.py --steps 1 --warmup 1 --workload c4 --no-reference-gpu > gpurun_out/ncu_nmr.log 2>&1; tail -2 gpurun_out/ncu_nmr.log
ls gpurun_out | head -50
