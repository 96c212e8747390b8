#!/bin/bash
# Round-2 session H: full GPU suite on the single-forward-kernel library, NMR (thread-per-face z-buffer, K9 with lane-serial
# in-scans) with occupancy A/B, ncu of the NMR kernels.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
for v in "" _k9m4 _k9m6; do
  B200R_LIB=$PWD/jrender_b200/lib/libb200raster$v.so timeout 600 python bench.py --steps 5 --warmup 3 --workload c4 --no-reference-gpu 2> gpurun_out/bench_c4$v.err | tee gpurun_out/bench_c4$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4$v', d['ms_per_step'], {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()})"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nmr_ -s 5 -c 5 -f -o gpurun_out/prof_nmr4 python bench.py --steps 1 --warmup 1 --workload c4 --no-reference-gpu > gpurun_out/ncu_nmr4.log 2>&1; tail -1 gpurun_out/ncu_nmr4.log
