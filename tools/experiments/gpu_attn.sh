#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 120 -p no:cacheprovider -x -k "attention" 2>&1 | tail -15
for at in 0 1; do
  MUGD_ATTN=$at timeout 300 python tools/profile_ops.py --B 4 2>&1 | grep -E "attention|sum of" 
done
MUGD_ATTN=1 timeout 300 python tools/profile_ops.py --B 8 --L 992 2>&1 | grep -E "attention|sum of"
MUGD_ATTN=0 timeout 300 python tools/profile_ops.py --B 8 --L 992 2>&1 | grep -E "attention|sum of"
