#!/bin/bash
# Round-2 session E: full suite on the cleaned-up library, TMA record staging A/B on the default forward kernel,
# ncu of the rewritten NMR edge-scan kernel, C1 / C4 benches.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
for lib in jrender_b200/lib/libb200raster.so jrender_b200/lib/libb200raster_tma.so; do
  for w in c3 c2 c5s; do
    B200R_LIB=$PWD/$lib timeout 300 python tools/ab_fwd2.py $w 2>&1 | grep -v Warning | tail -1
  done
done | tee gpurun_out/ab_tma.log
timeout 600 python bench.py --steps 60 --warmup 5 --workload c1 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; cut -c1-1200 gpurun_out/bench_c1.json; tail -2 gpurun_out/bench_c1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nmr_backward_pixel_map -s 1 -c 1 -f -o gpurun_out/prof_nmr_bwd2 python bench.py --steps 1 --warmup 1 --workload c4 --no-reference-gpu > gpurun_out/ncu_nmr2.log 2>&1; tail -1 gpurun_out/ncu_nmr2.log
timeout 900 python bench.py --steps 5 --warmup 3 --workload c4 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; cut -c1-300 gpurun_out/bench_c4.json
