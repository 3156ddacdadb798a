#!/bin/bash
for g in 1 2 4; do for m in 4 8; do
  echo "== groups=$g minb=$m"; DELORA_DENSE_GROUPS=$g DELORA_DENSE_MINB=$m python scripts/gpu_explore.py 8 2048 2>&1 | grep -E "icp_dense|full step"
done; done
