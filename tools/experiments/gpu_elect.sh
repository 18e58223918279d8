#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_ops.py -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -3
timeout 200 python tools/gemm_timeline.py 2>&1 | grep -A 10 "^== 1x1 256->256 M2048 bn128\|^== conv3 640->256 B64 bn256" | head -30
for wl in L512_B4_cfg5_S50 L512_B32_cfg5_S50 L992_B8_cfg5_S100; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/b.log 2>gpurun_out/b.err
  python -c "
import json;d=json.loads(open('gpurun_out/b.log').read());print('$wl', round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['family_ms_in_graph'], round(d['roofline']['achieved'],1))" || tail -3 gpurun_out/b.err
done
