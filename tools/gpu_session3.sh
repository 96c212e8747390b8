#!/bin/bash
# ncu --set full captures with source: C3 forward/backward and the C5 (demo2, silhouette, sigma 1e-4) forward/backward;
# then the parity suite + C3/C5 bench after the set_materialize_grads fix.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_softras -s 3 -c 3 -f -o gpurun_out/prof_c3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1; tail -2 gpurun_out/ncu_c3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_softras -s 1 -c 3 -f -o gpurun_out/prof_c5 python examples/demo2_deform.py --iters 2 --image-size 512 --batch-size 24 > gpurun_out/ncu_c5full.log 2>&1; tail -2 gpurun_out/ncu_c5full.log
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cut -c1-400 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
timeout 600 python bench.py --workload c5 --steps 60 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; cut -c1-1300 gpurun_out/bench_c5.json | tail -c 700
ls -la gpurun_out | head -30
