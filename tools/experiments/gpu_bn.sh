#!/bin/bash
mkdir -p gpurun_out
for bn in 0 128 256; do echo "== bench_gemm MUGD_TC_BN=$bn"; MUGD_TC_BN=$bn timeout 200 python tools/bench_gemm.py 2>&1 | tail -16 | cut -c1-150; done
for cfg in "MUGD_TC_BN=128" "MUGD_TC_NARROW=0.25" "MUGD_TC_NARROW=0.35"; do for wl in L512_B4_cfg5_S50 L512_B32_cfg5_S50; do
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/b.log 2>gpurun_out/b.err
  python -c "
import json;d=json.loads(open('gpurun_out/b.log').read());print('$cfg $wl', round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['family_ms_in_graph']['gemm'], d['roofline']['family_launches']['gemm'])" || tail -3 gpurun_out/b.err
done; done
