#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -3
timeout 200 python tools/bench_gemm.py 2>&1 | tail -16 | cut -c1-150 | grep -E "qkv|tiny|ff1 256|l3 1x1|big conv"
for wl in L512_B4_cfg5_S50 L512_B32_cfg5_S50 L992_B8_cfg5_S100; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/bench_$wl.log 2>gpurun_out/b.err
  python -c "
import json;d=json.loads(open('gpurun_out/bench_$wl.log').read());print('$wl', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['roofline']['family_ms_in_graph'], round(d['roofline']['achieved'],1))" || tail -3 gpurun_out/b.err
done
