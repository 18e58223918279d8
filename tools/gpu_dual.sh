#!/bin/bash
mkdir -p gpurun_out
P=/root/repo/mug_diffusion_b200
MUGD_LIB=$P/libmugd_dual.so timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_e2e.py -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -2
for lib in libmugd.so libmugd_dual.so; do
MUGD_LIB=$P/$lib timeout 200 python tools/bench_gemm.py 2>&1 | tail -16 | cut -c1-150 | grep -E "qkv|big conv|l2 conv3|tiny|l1 conv3"
for wl in L512_B4_cfg5_S50 L512_B32_cfg5_S50; do
  MUGD_LIB=$P/$lib timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/b.log 2>gpurun_out/b.err
  python -c "
import json;d=json.loads(open('gpurun_out/b.log').read());print('$lib $wl', round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['family_ms_in_graph']['gemm'], round(d['roofline']['achieved'],1))" || tail -3 gpurun_out/b.err
done; done
MUGD_TC_BN=128 MUGD_LIB=$P/libmugd_dual.so timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload L512_B32_cfg5_S50 > gpurun_out/b.log 2>gpurun_out/b.err
python -c "
import json;d=json.loads(open('gpurun_out/b.log').read());print('dual BN=128 forced B32', round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['family_ms_in_graph']['gemm'], round(d['roofline']['achieved'],1))"
