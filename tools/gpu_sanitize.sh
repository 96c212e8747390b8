#!/bin/bash
# compute-sanitizer on the kernels touched in the re-entry session (vector id stores, pixel-major top-K lists, lane-mask
# build, warp-aggregated backward, pre-raster stage, mesh losses)
mkdir -p gpurun_out
SEL="default_params_256 or non_multiple or topk_sizes or degenerate or vertex_textures or silhouette_mode or scheduling_choice or backward_modes"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_softras_gpu.py tests/test_preraster_gpu.py tests/test_mesh_loss_gpu.py -q -m gpu -x -k "$SEL or preraster or fused or project or transform or out_of_range or losses" > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_memcheck.log | tail -3
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_softras_gpu.py -q -m gpu -x -k "default_params_256 or silhouette_mode or topk_sizes" > gpurun_out/san_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/san_racecheck.log | tail -3
