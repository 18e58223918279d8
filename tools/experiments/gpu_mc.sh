#!/bin/bash
mkdir -p gpurun_out
for mc in 2 4; do
  MUGD_TC_MC=$mc timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -q --timeout 120 -p no:cacheprovider -x -k "multicast or conv3_same or linear" 2>&1 | tail -4
done
for g in auto tc_tf32; do for wl in L512_B4_cfg5_S50 L512_B32_cfg5_S50; do for mc in 0 2 4; do
  MUGD_TC_MC=$mc timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload $wl --gemm $g > gpurun_out/b.log 2>gpurun_out/b.err
  python -c "
import json;d=json.loads(open('gpurun_out/b.log').read());print('$g $wl MC=$mc', round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['family_ms_in_graph']['gemm'], round(d['roofline']['achieved'],1))" || tail -3 gpurun_out/b.err
done; done; done
MUGD_TC_MC=0 timeout 200 python tools/bench_gemm.py 2>&1 | tail -3
MUGD_TC_MC=2 timeout 200 python tools/bench_gemm.py 2>&1 | tail -3
