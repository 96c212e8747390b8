#!/bin/bash
# GPU session (re-entry): full parity suite incl. the fused pre-raster stage, C3 bench, C5 (demo2 loop) bench,
# demo2 at 64^2 eager/graph, ncu launch list of the C5 iteration.
mkdir -p gpurun_out
nvidia-smi -L; nproc
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cut -c1-900 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
timeout 600 python bench.py --workload c5 --steps 60 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; cat gpurun_out/bench_c5.json; tail -3 gpurun_out/bench_c5.err
timeout 300 python examples/demo2_deform.py --iters 200 2>&1 | tail -2 | tee gpurun_out/demo2_eager.log
timeout 300 python examples/demo2_deform.py --iters 200 --cuda-graph 2>&1 | tail -2 | tee gpurun_out/demo2_graph.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_c5.csv python examples/demo2_deform.py --iters 3 --image-size 512 --batch-size 120 > gpurun_out/ncu_c5.log 2>&1
ls gpurun_out
