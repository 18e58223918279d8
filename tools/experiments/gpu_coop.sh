#!/bin/bash
mkdir -p gpurun_out
MUGD_TC_COOP=2.0 timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_e2e.py -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -3
for c in "" 1.0 2.0 3.0; do for wl in L512_B4_cfg5_S50 L512_B32_cfg5_S50; do
  MUGD_TC_COOP=$c timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/b.log 2>gpurun_out/b.err
  python -c "
import json;d=json.loads(open('gpurun_out/b.log').read());print('coop=$c $wl', round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['family_ms_in_graph']['gemm'], d['roofline']['family_launches']['gemm'], d['config']['outputs_finite'])" || tail -3 gpurun_out/b.err
done; done
MUGD_TC_COOP=2.0 timeout 200 python tools/bench_gemm.py 2>&1 | tail -16 | cut -c1-150 | grep -E "l3 1x1|l0 conv3 128|l2 conv3|l1 ff2"
