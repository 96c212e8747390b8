mkdir -p gpurun_out
python -m oracle.make_ref_golden --bake-only 2>&1 | tail -8
cp gpurun_out/golden/ref_gpu_bake_n3mr_*.npz tests/golden/ 2>/dev/null
timeout 900 python -m pytest tests/test_bake_gpu.py -q -m gpu -x 2>&1 | tail -5
timeout 300 python -m pytest tests/test_golden.py -q -k bake 2>&1 | tail -3
