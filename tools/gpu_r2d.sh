#!/bin/bash
# Round-2 session D: full GPU suite after the AA / lighting / bake / cleanup work, bake golden, ncu of both forward kernels,
# secondary benches (C1 and C5 as configured, C4).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python -m oracle.make_ref_golden --bake-only 2>&1 | tail -2
timeout 600 python bench.py --steps 60 --warmup 5 --workload c1 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; cut -c1-900 gpurun_out/bench_c1.json; tail -2 gpurun_out/bench_c1.err
timeout 900 python bench.py --steps 30 --workload c5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; cut -c1-1500 gpurun_out/bench_c5.json; tail -2 gpurun_out/bench_c5.err
timeout 600 python bench.py --steps 5 --warmup 3 --workload c4 --no-reference-gpu > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; cut -c1-200 gpurun_out/bench_c4.json
for v in 1 2; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_softras_forward -s 1 -c 1 -f -o gpurun_out/prof_fwd_v$v python bench.py --steps 1 --warmup 1 --no-cpu-baseline --option softras_fwd_variant=$v > gpurun_out/ncu_fwd_v$v.log 2>&1; tail -1 gpurun_out/ncu_fwd_v$v.log
done
