timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mirror.py -q -m gpu -x 2>&1 | tail -4
timeout 200 python scripts/gpu_explore.py 2>&1 | grep -E "^projection|^normals|^icp_dense|^full step|^sum of|^icp_identityT|^icp_badT"
python -c "
import __graft_entry__ as g
g.smoke()" 2>&1 | tail -3
