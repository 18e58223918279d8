#!/bin/bash
mkdir -p gpurun_out
P=/root/repo/mug_diffusion_b200
for cfg in "MUGD_PDL=0" "MUGD_PDL=1" "MUGD_PDL=1 MUGD_LIB=$P/libmugd_late.so" "MUGD_PDL=1 MUGD_LIB=$P/libmugd_lateonly.so"; do for wl in L512_B4_cfg5_S50; do
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/b.log 2>gpurun_out/b.err
  python -c "
import json;d=json.loads(open('gpurun_out/b.log').read());print('$cfg $wl', round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['family_ms_in_graph'], 'finite', d['config']['outputs_finite'])" || tail -3 gpurun_out/b.err
done; done
MUGD_PDL=1 MUGD_LIB=$P/libmugd_late.so timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -2
