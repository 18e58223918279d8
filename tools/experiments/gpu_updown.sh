#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -q --timeout 60 -p no:cacheprovider -x 2>&1 | tail -8 > gpurun_out/pytest_tc.log; tail -8 gpurun_out/pytest_tc.log
MUGD_TC_BN=256 timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -q --timeout 60 -p no:cacheprovider -x 2>&1 | tail -4
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -5
timeout 200 python tools/bench_gemm.py 2>&1 | tail -17 | tee gpurun_out/bench_gemm.log
for v in "MUGD_TC_BN=0" "MUGD_TC_BN=128"; do
  env $v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/b.log 2>gpurun_out/b.err
  python -c "
import json;d=json.loads(open('gpurun_out/b.log').read());print('$v', round(d['value'],1), round(d['ms_per_step'],3), d['launches_per_step'], d['roofline']['family_ms_in_graph'])" || tail -3 gpurun_out/b.err
done
env MUGD_TC_BN=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload L512_B32_cfg5_S50 > gpurun_out/b32.log 2>gpurun_out/b.err
python -c "
import json;d=json.loads(open('gpurun_out/b32.log').read());print('B32 auto', round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['achieved'], d['roofline']['family_ms_in_graph'])" || tail -3 gpurun_out/b.err
