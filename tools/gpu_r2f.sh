#!/bin/bash
# Round-2 session F: NMR K9 (hoisted loads, 32-bit offsets) incl. occupancy A/B, SoftRas backward 4-level merge A/B, default bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nmr_gpu.py -q -m gpu > gpurun_out/pytest_nmr.log 2>&1; tail -2 gpurun_out/pytest_nmr.log
for v in "" _k9m5; do
  B200R_LIB=$PWD/jrender_b200/lib/libb200raster$v.so timeout 600 python bench.py --steps 5 --warmup 3 --workload c4 --no-reference-gpu 2> gpurun_out/bench_c4$v.err | tee gpurun_out/bench_c4$v.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4$v', d['ms_per_step'], {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()})"
done
for v in "" _merge2; do
  for w in c3 c2; do
    B200R_LIB=$PWD/jrender_b200/lib/libb200raster$v.so timeout 300 python tools/ab_fwd2.py $w 2>&1 | grep -v Warning | tail -1 | cut -c1-330
  done
done | tee gpurun_out/ab_merge.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cut -c1-400 gpurun_out/bench_c3.json; tail -2 gpurun_out/bench_c3.err
