#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe). Outputs in gpurun_out/.
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline ${BENCH_ARGS}"
# every launch of one DDIM step with its device time (skip session build + eager warm-up + graph steps)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-1800} -c ${NCU_COUNT:-600} --csv \
    --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_bench.log 2>&1
# the dominant kernel, full set, 3 launches
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-gemm_tc_kernel} -s 100 -c 3 -f -o gpurun_out/prof_gemm \
    $BENCH > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/ | tail -8
