#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_ops.py -m gpu -q --timeout 300 -p no:cacheprovider --durations=4 2>&1 | tail -22 > gpurun_out/pytest_tc.log
tail -22 gpurun_out/pytest_tc.log
timeout 300 python tools/bench_gemm.py 2>&1 | tail -17 | tee gpurun_out/bench_gemm.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>gpurun_out/bench.err; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
