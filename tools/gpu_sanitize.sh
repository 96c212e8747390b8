#!/bin/bash
# compute-sanitizer over the kernels written or rewritten in round 2: one-warp forward with the anti-aliasing epilogue,
# backward with the pooled-gradient prologue, NMR z-buffer / resolve / edge scans (shuffles, lane-serial scans, shared-memory
# face table), surface lighting, texture bakes (SoftRas and NMR), the pooled binning lists incl. the pool-exhausted path
# -- plus the round-1 selection.
mkdir -p gpurun_out
SEL="default_params_256 or non_multiple or topk_sizes or degenerate or vertex_textures or silhouette_mode or scheduling_choice or backward_modes or fused_antialiasing or coarse_list_pool or lists_span_chunks or large_sigma"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_softras_gpu.py tests/test_preraster_gpu.py tests/test_mesh_loss_gpu.py tests/test_lighting_gpu.py tests/test_bake_gpu.py tests/test_nmr_gpu.py -q -m gpu -x -k "$SEL or preraster or fused or project or transform or out_of_range or losses or lighting or bake_kernel or n3mr_bake or rgbad_sphere or output_subsets or texture_size_1 or overlapping_random or 3280_faces_256" > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_memcheck.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_softras_gpu.py tests/test_nmr_gpu.py -q -m gpu -x -k "default_params_256 or silhouette_mode or topk_sizes or fused_antialiasing or coarse_list_pool or rgbad_sphere or overlapping_random" > gpurun_out/san_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/san_racecheck.log | tail -3
