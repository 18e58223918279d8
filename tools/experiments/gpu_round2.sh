#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/pytest.log; tail -6 gpurun_out/pytest.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2>gpurun_out/bench.err; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --workload L512_B32_cfg5_S50 --no-cpu-baseline > gpurun_out/bench_b32.log 2>gpurun_out/bench_b32.err; cat gpurun_out/bench_b32.log; tail -3 gpurun_out/bench_b32.err
timeout 600 python bench.py --steps 20 --warmup 3 --workload L992_B8_cfg5_S100 --no-cpu-baseline > gpurun_out/bench_l992.log 2>gpurun_out/bench_l992.err; cat gpurun_out/bench_l992.log; tail -3 gpurun_out/bench_l992.err
