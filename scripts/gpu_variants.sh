#!/bin/bash
for m in 8 16 24 40 100000; do
  echo "== max_strips=$m"; DELORA_ICP_MAX_STRIPS=$m python scripts/gpu_explore.py 8 2048 2>&1 | grep -E "icp_dense|icp_identityT"
done
