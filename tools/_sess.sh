timeout 600 python -m pytest tests/test_bake_gpu.py -q -m gpu -x -k nmr_renderer 2>&1 | grep -E "AssertionError|passed|failed" | head
