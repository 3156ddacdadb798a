timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -4
timeout 200 python scripts/gpu_explore.py 2>&1 | grep -E "^icp_dense|^full step|^icp_identityT|^icp_badT"
timeout 300 python scripts/gpu_train_step.py 16 2>&1 | head -1
