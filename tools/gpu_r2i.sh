#!/bin/bash
# Round-2 session I: NMR with the pipelined K9 scan loop (occupancy A/B), expanded z-buffer pairs, 16-byte texture-gradient
# atomics; NMR parity suite; ncu of the NMR kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nmr_gpu.py -q -m gpu > gpurun_out/pytest_nmr.log 2>&1; tail -3 gpurun_out/pytest_nmr.log
for v in "" _k9m5 _k9m4; do
  B200R_LIB=$PWD/jrender_b200/lib/libb200raster$v.so timeout 600 python bench.py --steps 5 --warmup 3 --workload c4 --no-reference-gpu 2> gpurun_out/bench_c4$v.err | tee gpurun_out/bench_c4$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4$v', d['ms_per_step'], {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()})"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nmr_ -s 5 -c 5 -f -o gpurun_out/prof_nmr5 python bench.py --steps 1 --warmup 1 --workload c4 --no-reference-gpu > gpurun_out/ncu_nmr5.log 2>&1; tail -1 gpurun_out/ncu_nmr5.log
