#!/bin/bash
# Profiling pass of one build (gpurun command): whole-step launch lists with DRAM traffic for the bench workloads + one --set full
# capture of the representative ops of every kernel family.  Outputs under gpurun_out/<tag>/ ; summarise here with
# tools/summarize_step_ncu.py and tools/summarize_ncu_raw.py, commit the summaries under profiles/.
tag=${1:-prof}
out=gpurun_out/$tag
mkdir -p $out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for wl in L512_B4_cfg5_S50 L512_B32_cfg5_S50 L992_B8_cfg5_S100; do
  timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv --log-file $out/step_$wl.csv \
      python tools/ncu_step.py --workload $wl > $out/step_$wl.log 2>&1
  echo "step $wl rc=$? lines=$(wc -l < $out/step_$wl.csv)"
done
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -o $out/families python tools/ncu_families.py --B 4 --L 512 > $out/families.log 2>&1
echo "families rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -o $out/families_B32 python tools/ncu_families.py --B 32 --L 512 > $out/families_B32.log 2>&1
echo "families_B32 rc=$?"
ls -la $out
