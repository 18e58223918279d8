#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_e2e.py -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -3
timeout 200 python tools/bench_gemm.py 2>&1 | tail -16 | cut -c1-150
for c in "" "0.55,1.1,3" "0.55,1.1,2"; do for wl in L512_B4_cfg5_S50 L512_B32_cfg5_S50; do
  MUGD_TC_COST=$c timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/b.log 2>gpurun_out/b.err
  python -c "
import json;d=json.loads(open('gpurun_out/b.log').read());print('cost=$c $wl', round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['family_ms_in_graph']['gemm'], d['roofline']['family_launches']['gemm'])" || tail -3 gpurun_out/b.err
done; done
