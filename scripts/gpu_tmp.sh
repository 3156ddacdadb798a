timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mirror.py -q -m gpu -x 2>&1 | tail -5
for mult in 4 8 16; do
echo "== mult $mult"
DELORA_ICP_PEND_MULT=$mult timeout 200 python scripts/gpu_explore.py 2>&1 | grep -E "^icp_dense|^icp_identityT|^icp_badT|^full step"
done
for ms in 8 12 24; do
echo "== max_strips $ms"
DELORA_ICP_MAX_STRIPS=$ms timeout 200 python scripts/gpu_explore.py 2>&1 | grep -E "^icp_dense|^icp_identityT|^icp_badT|^full step"
done
