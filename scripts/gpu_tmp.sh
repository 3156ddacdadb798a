timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "normals or pipeline" 2>&1 | tail -3
timeout 200 python scripts/gpu_explore.py 2>&1 | grep -E "^projection|^normals|^icp_dense|^full step|^sum of"
