#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/eval_tf32.py 2>&1 | tail -22 | tee gpurun_out/eval_tf32.json
for wl in L512_B4_cfg5_S50 L512_B32_cfg5_S50; do
  MUGD_TC_TF32=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --workload $wl > gpurun_out/b.log 2>gpurun_out/b.err
  python -c "
import json;d=json.loads(open('gpurun_out/b.log').read());print('TF32 single pass $wl', round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['family_ms_in_graph']['gemm'], d['roofline']['achieved'])" || tail -3 gpurun_out/b.err
done
