bash tools/gpu_final.sh
out=gpurun_out/r02/prof3
mkdir -p $out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for wl in L512_B4_cfg5_S50 L512_B32_cfg5_S50 L992_B8_cfg5_S100; do
  timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --cache-control none --csv --log-file $out/step_$wl.csv \
      python tools/ncu_step.py --workload $wl > $out/step_$wl.log 2>&1
  echo "step $wl rc=$? lines=$(wc -l < $out/step_$wl.csv)"
done
