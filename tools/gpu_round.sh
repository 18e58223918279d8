#!/bin/bash
# full GPU verification + evidence for profiles/: tests, bench, launch list, per-family ncu --set full
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -5 | tee gpurun_out/pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.log 2>gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/families python tools/ncu_families.py > gpurun_out/families.txt 2>&1
tail -25 gpurun_out/families.txt
ncu -i gpurun_out/families.ncu-rep --page raw --csv > gpurun_out/families_raw.csv 2>/dev/null
ls -la gpurun_out/ | head -30
for wl in L512_B32_cfg5_S50 L992_B8_cfg5_S100; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/bench_$wl.log 2>gpurun_out/bench_$wl.err; tail -c 1500 gpurun_out/bench_$wl.log | head -c 400; echo
done
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
