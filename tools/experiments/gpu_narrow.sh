#!/bin/bash
mkdir -p gpurun_out
MUGD_TC_NARROW=0.4 timeout 600 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_e2e.py -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -3
for nk in "" 0.3 0.45 0.6; do for wl in L512_B4_cfg5_S50 L512_B32_cfg5_S50; do
  MUGD_TC_NARROW=$nk timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/b.log 2>gpurun_out/b.err
  python -c "
import json;d=json.loads(open('gpurun_out/b.log').read());print('narrow=$nk $wl', round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['family_ms_in_graph'], d['roofline']['family_launches']['gemm'])" || tail -3 gpurun_out/b.err
done; done
MUGD_TC_NARROW=0.45 timeout 300 python tools/profile_ops.py --B 4 > gpurun_out/ops_B4_narrow.txt 2>&1
timeout 300 python tools/profile_ops.py --B 4 > gpurun_out/ops_B4_base.txt 2>&1
