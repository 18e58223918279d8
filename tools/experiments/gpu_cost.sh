#!/bin/bash
mkdir -p gpurun_out
for c in "0.7,1.05,5" "0.55,1.1,5" "0.55,1.1,4" "0.55,1.1,6.5" "0.45,1.0,5" "0.55,1.3,5" "0.55,0.95,5" "0.6,1.1,8"; do for wl in L512_B4_cfg5_S50 L512_B32_cfg5_S50; do
  MUGD_TC_COST=$c timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/b.log 2>gpurun_out/b.err
  python -c "
import json;d=json.loads(open('gpurun_out/b.log').read());print('cost=$c $wl', round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['family_ms_in_graph']['gemm'], d['roofline']['family_launches']['gemm'])" || tail -3 gpurun_out/b.err
done; done
