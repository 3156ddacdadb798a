timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3
for th in 256 128; do for mult in 4 8; do
echo "== threads $th mult $mult"
DELORA_ICP_PEND_THREADS=$th DELORA_ICP_PEND_MULT=$mult timeout 200 python scripts/gpu_explore.py 2>&1 | grep -E "^icp_dense|^icp_identityT|^icp_badT|^full step"
done; done
