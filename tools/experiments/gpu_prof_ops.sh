#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/profile_ops.py --B 4 > gpurun_out/ops_B4.txt 2>gpurun_out/ops_B4.err || tail -5 gpurun_out/ops_B4.err
timeout 600 python tools/profile_ops.py --B 32 > gpurun_out/ops_B32.txt 2>gpurun_out/ops_B32.err || tail -5 gpurun_out/ops_B32.err
head -50 gpurun_out/ops_B4.txt
