#!/bin/bash
# end-of-round verification: GPU tests, smoke, bench lines (default + the two other single-GPU configs)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.log 2>gpurun_out/bench.err; tail -c 600 gpurun_out/bench.log; echo
for wl in L512_B32_cfg5_S50 L992_B8_cfg5_S100; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/bench_$wl.log 2>gpurun_out/bench_$wl.err
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err; tail -c 500 gpurun_out/bench_ref.log
