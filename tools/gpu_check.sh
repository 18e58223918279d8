#!/bin/bash
# Standard GPU validation of one build, meant to be the command of a gpurun call:
#   gpurun --timeout 900 -- 'bash tools/gpu_check.sh <tag> [pytest-args...]'
# Runs the -m gpu tests and the default bench line; everything lands under gpurun_out/<tag>/ (merged back by gpurun).
tag=${1:-check}; shift
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q "$@" > $out/pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $out/pytest.log
timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $out/bench.json 2> $out/bench.err
echo "bench rc=$?"; tail -2 $out/bench.err
python - <<PY
import json
try:
    d = json.load(open("$out/bench.json"))
    r = d["roofline"]
    print("steps/s", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "launches/step", d["launches_per_step"])
    print("family ms", r["family_ms_in_graph"], r["family_launches"])
    for k, w in d.get("secondary", {}).get("workloads", {}).items():
        print(k, round(w["value"], 1), "e2e", round(w["e2e"]["value"], 1), "frac3xtf32", round(w["roofline"].get("frac_of_3xtf32_ceiling", 0), 3), w["roofline"]["family_ms_in_graph"])
except Exception as e:
    print("no bench line:", e)
PY
