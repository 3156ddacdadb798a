timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_mirror.py -q -m gpu -x 2>&1 | tail -6
timeout 300 python scripts/gpu_train_step.py 16 2>&1 | tail -3
