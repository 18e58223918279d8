#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --workload L512_B32_cfg5_S50 --no-cpu-baseline > gpurun_out/bench_b32.log 2>gpurun_out/bench_b32.err; cat gpurun_out/bench_b32.log; tail -3 gpurun_out/bench_b32.err
timeout 600 python bench.py --steps 20 --warmup 3 --workload L992_B8_cfg5_S100 --no-cpu-baseline > gpurun_out/bench_l992.log 2>gpurun_out/bench_l992.err; cat gpurun_out/bench_l992.log; tail -3 gpurun_out/bench_l992.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_g2.log 2>gpurun_out/bench_g2.err; cat gpurun_out/bench_g2.log; tail -5 gpurun_out/bench_g2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/bench_g2ref.log 2>gpurun_out/bench_g2ref.err; cat gpurun_out/bench_g2ref.log | cut -c1-300; tail -3 gpurun_out/bench_g2ref.err
