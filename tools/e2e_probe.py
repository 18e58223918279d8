"""Where does a request spend host time?  cProfile of one warmed sampler.sample + decode request (bench.py's e2e leg)."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mug_diffusion_b200 import synth  # noqa: E402
from mug_diffusion_b200.config import ModelConfig  # noqa: E402
from mug_diffusion_b200.sampler import DDIMSampler, MugDiffusionB200  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "L512_B32_cfg5_S50"
wl = bench.WORKLOADS[name]
L, B = wl["L"], wl["B"]
dev = torch.device("cuda:0")
model = MugDiffusionB200(synth.synthetic_state_dict(L), ModelConfig(), z_length=L, device=dev)
sampler = DDIMSampler(model)
inp = bench.make_inputs(wl, 0)
host = dict(x_T=inp["x_T"].pin_memory(), c=inp["c"].pin_memory(), uc=inp["uc"].pin_memory(), w=[w.pin_memory() for w in inp["w"]])


def request(K=20):
    c = host["c"].to(dev, non_blocking=True)
    uc = host["uc"].to(dev, non_blocking=True)
    w = [t.to(dev, non_blocking=True) for t in host["w"]]
    xT = host["x_T"].to(dev, non_blocking=True)
    z, _ = sampler.sample(S=K, c=c, w=w, batch_size=B, shape=(16, L), verbose=False, x_T=xT, eta=0.0,
                          unconditional_guidance_scale=wl["scale"], unconditional_conditioning=uc, tqdm_class=bench._NoBar)
    logits = model.model.decode(z)
    return logits.to("cpu")


request()
torch.cuda.synchronize()
t0 = time.perf_counter()
request()
torch.cuda.synchronize()
print(f"{name}: request {1e3 * (time.perf_counter() - t0):.1f} ms for 20 steps")
pr = cProfile.Profile()
pr.enable()
request()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
