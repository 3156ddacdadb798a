timeout 900 python -m pytest tests/test_gpu_encoder.py -q -m gpu -x 2>&1 | tail -6
timeout 300 python scripts/gpu_encoder_bwd_test.py 16 2048 2>&1 | tail -4
