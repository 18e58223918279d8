"""Characterise the opt-in single-pass TF32 mode against the fp32-accurate default (run on the GPU box)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_cases as gc  # noqa: E402
from mug_diffusion_b200 import synth  # noqa: E402
from mug_diffusion_b200.sampler import DDIMSampler, MugDiffusionB200  # noqa: E402
from oracle import mug_oracle as orc  # noqa: E402


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).abs().max() / b.abs().max())


out = {}
for impl in ("tc", "tc_tf32"):
    r = {}
    m = MugDiffusionB200.from_state_dict(synth.synthetic_state_dict(512), z_length=512, gemm_impl=impl)
    inp = synth.synthetic_inputs(2, 512)
    eps = m.model.forward(inp["x_T"].cuda(), torch.tensor([501, 21]).cuda(), inp["c"].cuda(), [w.cuda() for w in inp["w"]])
    r["unet_eval_rel_err"] = rel(eps, gc.load_golden(os.path.join(ROOT, "tests/golden/unet_L512_B2.npz"))["eps"])
    inp = synth.synthetic_inputs(1, 512)
    z, _ = DDIMSampler(m).sample(S=50, c=inp["c"].cuda(), w=[w.cuda() for w in inp["w"]], batch_size=1, verbose=False, x_T=inp["x_T"].cuda(),
                                 unconditional_guidance_scale=5.0, unconditional_conditioning=inp["uc"].cuda(), shape=(16, 512))
    g = gc.load_golden(os.path.join(ROOT, "tests/golden/ddim_L512_B1_S50_cfg5.npz"))
    lg = m.model.decode(z)
    r["ddim50_z_rel_err"] = rel(z, g["z"])
    r["ddim50_logits_rel_err"] = rel(lg, g["logits"])
    flips = orc.notes_from_logits(lg.cpu()) != orc.notes_from_logits(g["logits"])
    r["note_decisions_flipped"] = int(flips.sum())
    r["note_decisions_total"] = int(flips.numel())
    out[impl] = r
    del m
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
