#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -q -s --timeout 120 -p no:cacheprovider -x 2>&1 | tail -60 > gpurun_out/pytest_tc.log
tail -45 gpurun_out/pytest_tc.log
