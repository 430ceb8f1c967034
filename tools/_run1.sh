timeout 600 python -m pytest tests/test_gpu_core.py -m gpu -q -x -k "oracle_pipeline" 2>&1 | tail -15
