#!/bin/bash
P=/root/repo/mug_diffusion_b200
echo "== base lib, batch32 test"; timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q --timeout 300 -p no:cacheprovider -k batch32 2>&1 | grep -E "assert|Error|passed|failed" | head -8
echo "== base lib forced BN=256 gemm tests"; MUGD_TC_BN=256 timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -q --timeout 120 -p no:cacheprovider 2>&1 | grep -E "^FAILED|passed|failed|assert " | head -12
echo "== dec lib forced BN=256 gemm tests"; MUGD_LIB=$P/libmugd_dec.so MUGD_TC_BN=256 timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -q --timeout 120 -p no:cacheprovider 2>&1 | grep -E "^FAILED|passed|failed|assert " | head -12
