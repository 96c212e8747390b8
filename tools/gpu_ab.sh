#!/bin/bash
# parity of the product build (quick subset) + interleaved in-process A/B of every lib/libb200raster*.so
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_softras_gpu.py tests/test_golden.py tests/test_exact_math_gpu.py -q -m gpu -x 2>&1 | tail -2
LIBS=$(ls $PWD/jrender_b200/lib/libb200raster*.so)
for w in ${AB_WORKLOADS:-c3 c5}; do timeout 600 python tools/ab_interleaved.py $w $LIBS 2>&1 | tail -1 | tee -a gpurun_out/ab_interleaved.log; done
