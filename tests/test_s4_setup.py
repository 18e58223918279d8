"""Host-side C~ lengthening (mug_diffusion_b200/s4_setup.py) against the reference's own in-place mutation
(tests/golden/s4_lengthen.npz from tools/make_goldens.py).  CPU only."""
import os

import pytest
import torch

import golden_cases as gc
from mug_diffusion_b200 import s4_setup, synth
from oracle import mug_oracle as orc

PRE = "model.unet_model.input_blocks.2.1.s4_model.kernel.kernel."


@pytest.mark.parametrize("tag,L_state,L_req", [("double", 48, 96), ("double2", 24, 96), ("init", 0, 96)])
def test_lengthen_matches_reference(tag, L_state, L_req, golden_dir):
    gold = gc.load_golden(os.path.join(golden_dir, "s4_lengthen.npz"))
    sd = synth.synthetic_state_dict(max(L_state, 8) if L_state else 96, decoder=False)
    params = {n: sd[PRE + n] for n in ("C", "log_dt", "P", "inv_w_real", "w_imag")}
    C_new, L_new = s4_setup.lengthen(params, L_state, L_req)
    assert L_new == int(gold[tag + ".L"][0]) and L_new >= L_req
    ref_C = gold[tag + ".C"]
    assert float((C_new - ref_C).abs().max() / ref_C.abs().max()) < 2e-5
    # and the kernel generated from the lengthened C~ equals the reference's kernel
    sd2 = dict(sd)
    sd2[PRE + "C"] = C_new
    sd2[PRE + "L"] = torch.tensor(L_new)
    k = orc.s4_nplr_kernel(sd2, PRE, L_req)
    assert float((k - gold[tag + ".K"]).abs().max() / gold[tag + ".K"].abs().max()) < 5e-5


def test_no_change_when_long_enough():
    sd = synth.synthetic_state_dict(96, decoder=False)
    params = {n: sd[PRE + n] for n in ("C", "log_dt", "P", "inv_w_real", "w_imag")}
    C_new, L_new = s4_setup.lengthen(params, 96, 64)
    assert L_new == 96 and torch.equal(C_new, params["C"])
