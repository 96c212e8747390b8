#!/bin/bash
# Round-2 session C: full GPU suite (incl. fused lighting, bake, C1), two-phase forward A/B (bit-identity + kernel times),
# sanitizer on the new kernels, reference-kernel golden for the texture bake.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python -m oracle.make_ref_golden --bake-only 2>&1 | tail -2
for lib in jrender_b200/lib/libb200raster*.so; do
  for w in c3 c2 c5s; do
    B200R_LIB=$PWD/$lib timeout 300 python tools/ab_fwd2.py $w 2>&1 | grep -v Warning | tail -1
  done
done | tee gpurun_out/ab_fwd2.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_softras_gpu.py -q -m gpu -x -k "every_forward_configuration and 1-2-" > gpurun_out/san_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/san_racecheck.log | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_softras_gpu.py tests/test_nmr_gpu.py tests/test_lighting_gpu.py tests/test_bake_gpu.py -q -m gpu -x -k "(every_forward_configuration and 1-2-) or rgbad_sphere or output_subsets or lighting or bake_kernel" > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_memcheck.log | tail -3
