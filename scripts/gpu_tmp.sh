timeout 600 python -m pytest tests -q -m gpu -x -s 2>&1 | grep -E "normals_|passed|failed|Error|assert" | head -40
timeout 200 python scripts/gpu_explore.py 2>&1 | grep -E "^projection|^normals|^icp_dense|^full step|^sum of"
