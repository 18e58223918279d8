"""Would two half-batch U-Net evals on two streams beat one full-batch eval?  (B=4 CFG: one Beff=8 graph vs two Beff=4 graphs
replayed concurrently.)  Each half gets its own engine so that split-K workspaces are not shared."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mug_diffusion_b200 import synth  # noqa: E402
from mug_diffusion_b200.config import ModelConfig  # noqa: E402
from mug_diffusion_b200.sampler import MugDiffusionB200  # noqa: E402

L = 512
dev = torch.device("cuda:0")
sd = synth.synthetic_state_dict(L)


def make(Beff):
    m = MugDiffusionB200(sd, ModelConfig(), z_length=L, device=dev)
    s = m.engine.session(Beff, L, per_sample_t=False)
    s.eval(graph=True)
    return m, s


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for total in (8, 16):
    m8, s8 = make(total)
    t_full = timeit(lambda: s8.eval(graph=True))
    for parts in (2, 4):
        ms_ = [make(total // parts) for _ in range(parts)]
        streams = [torch.cuda.Stream() for _ in range(parts)]

        def both():
            cur = torch.cuda.current_stream()
            for st, (_, s) in zip(streams, ms_):
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    s.eval(graph=True)
            for st in streams:
                cur.wait_stream(st)

        t_par = timeit(both)
        t_one = timeit(lambda: ms_[0][1].eval(graph=True))
        print(f"Beff={total}: one graph {t_full:.3f} ms | {parts} x Beff={total//parts} concurrent {t_par:.3f} ms | a single Beff={total//parts} graph alone {t_one:.3f} ms", flush=True)
        del ms_
    del m8, s8
    torch.cuda.empty_cache()
