#!/bin/bash
# one GPU-box visit: parity tests, smoke, short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider ${PYTEST_ARGS} 2>&1 | tail -${PYTEST_TAIL:-60} > gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps ${BENCH_STEPS:-50} --warmup 5 ${BENCH_ARGS} > gpurun_out/bench.log 2>gpurun_out/bench.err
tail -${PYTEST_TAIL:-60} gpurun_out/pytest.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
