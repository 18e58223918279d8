"""Import shims that let the UNMODIFIED reference (/root/reference) run on CPU in this container.

Test/fixture infrastructure only: used by tools/make_goldens.py (here, where /root/reference exists)
to pin oracle/ against the reference itself.  Nothing in the product path or on the GPU box imports it.

The four missing third-party packages are stubbed:
  pytorch_lightning (LightningModule -> nn.Module), opt_einsum (contract -> torch.einsum),
  omegaconf.listconfig.ListConfig, and import-time-only audio libs.
"""
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("MUG_REFERENCE_ROOT", "/root/reference")


def install_shims():
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            @property
            def device(self):
                try:
                    return next(self.parameters()).device
                except StopIteration:
                    return torch.device("cpu")

            def log(self, *a, **k):
                pass

            def log_dict(self, *a, **k):
                pass

        pl.LightningModule = LightningModule
        pl.Callback = object
        util = types.ModuleType("pytorch_lightning.utilities")
        util.rank_zero_only = lambda f: f
        dist = types.ModuleType("pytorch_lightning.utilities.distributed")
        dist.rank_zero_only = lambda f: f
        pl.utilities = util
        sys.modules["pytorch_lightning"] = pl
        sys.modules["pytorch_lightning.utilities"] = util
        sys.modules["pytorch_lightning.utilities.distributed"] = dist
    if "opt_einsum" not in sys.modules:
        oe = types.ModuleType("opt_einsum")
        oe.contract = lambda expr, *ops, **kw: torch.einsum(expr, *ops)

        def contract_expression(expr, *shapes, **kw):
            return lambda *ops: torch.einsum(expr, *ops)

        oe.contract_expression = contract_expression
        sys.modules["opt_einsum"] = oe
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")
        lc = types.ModuleType("omegaconf.listconfig")

        class ListConfig(list):
            pass

        lc.ListConfig = ListConfig
        oc.listconfig = lc
        oc.ListConfig = ListConfig
        sys.modules["omegaconf"] = oc
        sys.modules["omegaconf.listconfig"] = lc
    for name in ("audioread", "audioread.ffdec", "soundfile", "librosa", "torchsummary"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


def load_reference_model(yaml_rel="configs/mug/mug_diffusion.yaml", z_length=None):
    """Instantiate the reference DDPM (random init) from its shipped yaml. CPU, eval mode."""
    import yaml

    install_shims()
    cwd = os.getcwd()
    os.chdir(REF_ROOT)  # cond_stage yaml path is relative (mug_diffusion.yaml:72)
    try:
        with open(os.path.join(REF_ROOT, yaml_rel)) as f:
            cfg = yaml.safe_load(f)
        params = cfg["model"]["params"]
        params["ckpt_path"] = None
        params.pop("scheduler_config", None)
        params["wave_stage_config"]["params"]["use_checkpoint"] = False
        if z_length is not None:
            params["z_length"] = z_length
        from mug.util import instantiate_from_config

        model = instantiate_from_config(cfg["model"])
    finally:
        os.chdir(cwd)
    model.eval()
    return model, cfg
