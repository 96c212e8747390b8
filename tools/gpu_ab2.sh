#!/bin/bash
mkdir -p gpurun_out
P=$PWD/jrender_b200/lib/libb200raster.so
for w in c2 c5 c3; do timeout 600 python tools/ab_interleaved.py $w "$P#softras_bwd_variant=1" "$P#softras_bwd_variant=0" 2>&1 | tail -1 | tee -a gpurun_out/ab_bwd_variant.log; done
LIBS=$(ls $PWD/jrender_b200/lib/libb200raster*.so)
for w in c3 c5; do timeout 600 python tools/ab_interleaved.py $w $LIBS 2>&1 | tail -1 | tee -a gpurun_out/ab_interleaved.log; done
