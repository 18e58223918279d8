#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -q --timeout 60 -p no:cacheprovider -x 2>&1 | tail -5 > gpurun_out/pytest_tc.log; tail -5 gpurun_out/pytest_tc.log
if grep -q "passed" gpurun_out/pytest_tc.log && ! grep -q "failed" gpurun_out/pytest_tc.log; then
  for v in 0 4 8 16; do
    MUGD_TC_INKERNEL=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/b.log 2>gpurun_out/b.err
    python -c "
import json;d=json.loads(open('gpurun_out/b.log').read());print('INKERNEL=$v', round(d['value'],1), round(d['ms_per_step'],3), d['launches_per_step'], d['roofline']['family_ms_in_graph']['gemm'], d['roofline']['achieved'])" || tail -3 gpurun_out/b.err
  done
  timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -4
fi
