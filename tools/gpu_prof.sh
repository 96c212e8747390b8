#!/bin/bash
# final-state captures: ncu launch list of the bench step + ncu --set full (with source) of the raster kernels, C3 and C5
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_softras -s 3 -c 3 -f -o gpurun_out/prof_c3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c3.log 2>&1; tail -1 gpurun_out/ncu_c3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_softras -s 1 -c 3 -f -o gpurun_out/prof_c5 python examples/demo2_deform.py --iters 2 --image-size 512 --batch-size 24 > gpurun_out/ncu_c5full.log 2>&1; tail -1 gpurun_out/ncu_c5full.log
ls -la gpurun_out | grep -E "prof|launches"
