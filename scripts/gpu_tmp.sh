timeout 900 python -m pytest tests/test_gpu_mirror.py -q -m gpu -x 2>&1 | tail -15
