mkdir -p gpurun_out/r02/ab
for v in base leader base2; do
  if [ "$v" = "leader" ]; then export MUGD_LIB=mug_diffusion_b200/libmugd_leader.so; else unset MUGD_LIB; fi
  timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02/ab/$v.json 2> gpurun_out/r02/ab/$v.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r02/ab/$v.json"))
print("$v", round(d["value"],1), d["roofline"]["family_ms_in_graph"]["gemm"], [(k, round(w["value"],1), w["roofline"]["family_ms_in_graph"]["gemm"]) for k,w in d["secondary"]["workloads"].items()])
PY
done
if [ -n "$1" ]; then MUGD_LIB=mug_diffusion_b200/libmugd_leader.so timeout 300 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_fusion.py -x -q 2>&1 | tail -2; fi
