"""The drop-in constructor reads the architecture off a live reference DDPM.  Runs only where the reference tree exists
(the build container); the GPU box has no /root/reference, so the test skips there."""
import os
import sys

import pytest

REF = os.environ.get("MUG_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "mug")), reason="reference tree not present")


def test_config_from_reference_matches_shipped_yaml():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import ref_shim
    from mug_diffusion_b200 import netspec
    from mug_diffusion_b200.config import ModelConfig
    from mug_diffusion_b200.sampler import MugDiffusionB200

    model, _ = ref_shim.load_reference_model()
    sd, cfg = MugDiffusionB200.config_from_reference(model)
    want = ModelConfig()
    assert cfg.unet == want.unet
    assert cfg.decoder == want.decoder
    assert (cfg.z_channels, cfg.timesteps, cfg.linear_start, cfg.linear_end) == (16, 1000, 1e-4, 2e-2)
    # every tensor the packer needs is present in the reference state_dict under the names netspec generates
    need = {**netspec.unet_param_specs(cfg.unet), **netspec.decoder_param_specs(cfg.decoder)}
    assert all(k in sd and tuple(sd[k].shape) == tuple(shape) for k, (shape, _) in need.items())
