#!/bin/bash
# interleaved in-process A/B of every lib/libb200raster*.so (min / median / max over rounds)
mkdir -p gpurun_out
LIBS=$(ls $PWD/jrender_b200/lib/libb200raster*.so)
for w in c3 c5; do timeout 600 python tools/ab_interleaved.py $w $LIBS 2>&1 | tail -1 | tee -a gpurun_out/ab_interleaved.log; done
