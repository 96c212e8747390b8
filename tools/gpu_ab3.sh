#!/bin/bash
P=$PWD/jrender_b200/lib/libb200raster.so
for w in c2 c5; do timeout 600 python tools/ab_interleaved.py $w $P "$P#softras_fwd_variant=1" "$P#softras_bwd_variant=1" 2>&1 | tail -1; done
