#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/pytest.log
tail -12 gpurun_out/pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pdl1.log 2>gpurun_out/bench.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_pdl1.log').read());print('PDL on :',d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['eager_ms_by_kernel'])"; tail -3 gpurun_out/bench.err
MUGD_PDL=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pdl0.log 2>gpurun_out/bench.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_pdl0.log').read());print('PDL off:',d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['eager_ms_by_kernel'])"; tail -3 gpurun_out/bench.err
