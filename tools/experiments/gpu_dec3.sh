#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_e2e.py -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -3
MUGD_TC_BN=256 timeout 200 python tools/bench_gemm.py 2>&1 | tail -16 | cut -c1-150 | grep -E "big conv|ff1|qkv"
for cfg in "MUGD_TC_COST=0.55,1.1,4" "MUGD_TC_COST=0.55,0.9,4" "MUGD_TC_COST=0.55,0.8,4"; do for wl in L512_B4_cfg5_S50 L512_B32_cfg5_S50 L992_B8_cfg5_S100; do
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/b.log 2>gpurun_out/b.err
  python -c "
import json;d=json.loads(open('gpurun_out/b.log').read());print('$cfg $wl', round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['family_ms_in_graph']['gemm'], round(d['roofline']['achieved'],1))" || tail -3 gpurun_out/b.err
done; done
