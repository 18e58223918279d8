#!/bin/bash
# planner cost-model sweep (us per k-step of a 128- / 256-wide tile, us per split-K round trip): bench.py --no-secondary per point
mkdir -p gpurun_out/r02/sweep
for c in "0.65,0.9,3.0" "0.75,0.9,3.0" "0.55,0.9,2.0" "0.55,0.9,1.5" "0.65,0.9,2.0" "0.85,0.9,3.0"; do
  MUGD_TC_COST=$c timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/r02/sweep/b_$c.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/r02/sweep/b_$c.json")); r=d["roofline"]
print("cost $c", round(d["value"],1), "steps/s", round(d["ms_per_step"],3), "ms  gemm", r["family_ms_in_graph"]["gemm"], "launches", d["launches_per_step"])
PY
done
