#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/pytest.log; tail -15 gpurun_out/pytest.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/b.log 2>gpurun_out/b.err
python -c "
import json;d=json.loads(open('gpurun_out/b.log').read());print('B4', round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['achieved'], d['roofline']['family_ms_in_graph'])" || tail -3 gpurun_out/b.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload L992_B8_cfg5_S100 > gpurun_out/b992.log 2>gpurun_out/b.err
python -c "
import json;d=json.loads(open('gpurun_out/b992.log').read());print('L992', round(d['value'],1), round(d['ms_per_step'],3), d['roofline']['achieved'], d['roofline']['family_ms_in_graph'])" || tail -3 gpurun_out/b.err
