#!/bin/bash
mkdir -p gpurun_out
MUGD_TC_BATCH=1 timeout 900 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -2
for bt in 0 1; do
MUGD_TC_BATCH=$bt timeout 200 python tools/bench_gemm.py 2>&1 | tail -16 | cut -c1-150 | grep -E "qkv|big conv|l2 conv3|tiny"
for wl in L512_B4_cfg5_S50 L512_B32_cfg5_S50; do
  MUGD_TC_BATCH=$bt timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/b.log 2>gpurun_out/b.err
  python -c "
import json;d=json.loads(open('gpurun_out/b.log').read());print('batch=$bt $wl', round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['family_ms_in_graph']['gemm'], round(d['roofline']['achieved'],1))" || tail -3 gpurun_out/b.err
done; done
