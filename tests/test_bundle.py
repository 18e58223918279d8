"""The C boundary beyond Python: a compiled request saved as plan files + region contents, loaded and run by a plain C
host (examples/host_c/sample_host.c: gcc + CUDA runtime + libmugd.so, no torch) -- must reproduce the Python run."""
import os
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_DIR = os.path.join(ROOT, "examples", "host_c")


def test_c_host_builds_against_the_public_header():
    """gcc compiles and links the C host against include/mugd.h + libmugd.so (no GPU needed to build)"""
    from mug_diffusion_b200 import build
    build.build()
    r = subprocess.run(["make", "-C", HOST_DIR, "-B"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.path.exists(os.path.join(HOST_DIR, "sample_host"))


def test_plan_symbols_exported():
    import ctypes as C
    from mug_diffusion_b200 import lib as L_
    lib = L_.load()
    for s in ("mugd_sample", "mugd_plan_save", "mugd_plan_load", "mugd_plan_ops", "mugd_plan_regions"):
        assert hasattr(lib, s)


@pytest.mark.gpu
def test_c_host_reproduces_the_python_request(tmp_path):
    """export a CFG request (z_length 96, 10 DDIM steps + decode), run it with the C host: logits equal the Python run (same kernels,
    same plans; 1e-5 covers the summation order of the fp64 row-moment atomics) -- and the Python run equals sampler.sample()"""
    from mug_diffusion_b200 import synth
    from mug_diffusion_b200.bundle import export_bundle
    from mug_diffusion_b200.sampler import DDIMSampler, MugDiffusionB200
    L, B, S = 96, 2, 10
    model = MugDiffusionB200.from_state_dict(synth.synthetic_state_dict(L), z_length=L)
    inp = synth.synthetic_inputs(B, L)
    out = str(tmp_path / "bundle")
    res = export_bundle(model, inp, S, 5.0, out)
    z, _ = DDIMSampler(model).sample(S=S, c=inp["c"].cuda(), w=[w.cuda() for w in inp["w"]], batch_size=B, verbose=False, x_T=inp["x_T"].cuda(),
                                     eta=0.0, shape=(16, L), unconditional_guidance_scale=5.0, unconditional_conditioning=inp["uc"].cuda())
    logits = model.model.decode(z)
    assert float((res["z"] - z).abs().max() / z.abs().max()) < 1e-5
    assert float((res["logits"] - logits).abs().max() / logits.abs().max()) < 1e-5
    del model
    torch.cuda.empty_cache()
    subprocess.run(["make", "-C", HOST_DIR], check=True, capture_output=True)
    r = subprocess.run([os.path.join(HOST_DIR, "sample_host"), out], capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count(" OK") == 2 and "sampled 10 DDIM steps" in r.stdout
