"""One whole U-Net evaluation of a bench workload (every launch of the plan, in plan order) between cudaProfilerStart/Stop, so that

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --cache-control none --csv --log-file gpurun_out/<tag>/step_<workload>.csv \
        python tools/ncu_step.py --workload L512_B4_cfg5_S50

captures the per-launch time and DRAM traffic of exactly one step (the launch plan run eagerly; the CUDA graph replays the same
kernels with the same arguments).  --cache-control none keeps L2 warm across launches like the real step does; the evaluation before
the profiled one warms the weights that fit.  tools/summarize_step_ncu.py turns the CSV into profiles/r02_step_*.{md,json}."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS, make_inputs  # noqa: E402
from mug_diffusion_b200 import synth  # noqa: E402
from mug_diffusion_b200.sampler import MugDiffusionB200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="L512_B4_cfg5_S50", choices=list(WORKLOADS))
    a = ap.parse_args()
    wl = WORKLOADS[a.workload]
    L, B = wl["L"], wl["B"]
    cfg_on = wl["scale"] != 1.0
    Beff = 2 * B if cfg_on else B
    dev = torch.device("cuda:0")
    m = MugDiffusionB200(synth.synthetic_state_dict(L), z_length=L, device=dev)
    inp = make_inputs(wl, 0)
    sess = m.engine.session(Beff, L, per_sample_t=False)
    sess.set_timestep_table([981, 961, 941])
    sess.set_context([inp["uc"].to(dev), inp["c"].to(dev)] if cfg_on else inp["c"].to(dev))
    sess.set_audio([w.to(dev) for w in inp["w"]], dup=cfg_on)
    sess.load_x(inp["x_T"].to(dev), dup=cfg_on)
    sess.set_step(0)
    sess.eval(graph=False)
    sess.eval(graph=False)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    sess.eval(graph=False)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("ops", sess.plan.n_ops, "launches", sess.plan.launches)


if __name__ == "__main__":
    main()
