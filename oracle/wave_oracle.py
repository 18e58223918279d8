"""CPU oracle of the audio encoder  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE (see mug_oracle.py header).

Functional torch-fp32 restatement of ``MelspectrogramScaleEncoder1D.forward`` (mug/cond/wave.py:398-467) over a flat
state_dict with the reference's key names (prefix ``model.wave_model.``).  Pinned by tests/golden/wave_T6144_B2.npz,
produced by the unmodified reference (tools/make_goldens.py --only wave).
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

from .mug_oracle import conv1d, cross_attention, downsample, group_norm, linear

WAVE_PREFIX = "model.wave_model."
DEFAULT_WAVE = dict(n_freq=128, middle_channels=128, attention_resolutions=(128, 256, 512), num_res_blocks=2, num_heads=8,
                    num_groups=32, channel_mult=(1, 1, 1, 1, 2, 2, 2, 4, 4, 4))


def dilated_resnet_block(p, pre: str, x: torch.Tensor, groups: int, dil: Sequence[int]) -> torch.Tensor:
    """ResnetBlock with dilations (d1, d2), temb_channels=0  -- mug/model/models.py:94-159, wave.py:425-433."""
    h = F.silu(group_norm(p, pre + "norm1.", x, groups))
    h = F.conv1d(h, p[pre + "conv1.weight"], p[pre + "conv1.bias"], padding=dil[0], dilation=dil[0])
    h = F.silu(group_norm(p, pre + "norm2.", h, groups))
    h = F.conv1d(h, p[pre + "conv2.weight"], p[pre + "conv2.bias"], padding=dil[1], dilation=dil[1])
    if pre + "nin_shortcut.weight" in p:
        x = conv1d(p, pre + "nin_shortcut.", x)
    return x + h


def self_transformer(p, pre: str, x: torch.Tensor, heads: int) -> torch.Tensor:
    """ContextualTransformer with context=None: GN32 -> 1x1 -> [LN-attn1-+ ; LN-attn2(self)-+ ; LN-GEGLU-+] -> 1x1 -> +x
    -- mug/model/attention.py:186-199, 147-151 (attn2 falls back to self-attention, :94)."""
    C = x.shape[1]
    h = conv1d(p, pre + "proj_in.", group_norm(p, pre + "norm.", x, 32)).transpose(1, 2)
    t = pre + "transformer_blocks.0."

    def ln(n, y):
        return F.layer_norm(y, (C,), p[t + n + ".weight"], p[t + n + ".bias"], eps=1e-5)

    h = cross_attention(p, t + "attn1.", ln("norm1", h), None, heads) + h
    h = cross_attention(p, t + "attn2.", ln("norm2", h), None, heads) + h
    a, g = linear(p, t + "ff.net.0.proj.", ln("norm3", h)).chunk(2, dim=-1)
    h = linear(p, t + "ff.net.2.", a * F.gelu(g)) + h
    return conv1d(p, pre + "proj_out.", h.transpose(1, 2)) + x


def wave_forward(p: Dict[str, torch.Tensor], mel: torch.Tensor, cfg: dict = DEFAULT_WAVE, prefix: str = WAVE_PREFIX) -> List[torch.Tensor]:
    """mel [B,128,T] -> list of the 10 level outputs  -- mug/cond/wave.py:453-467."""
    h = conv1d(p, prefix + "conv_in.", mel, padding=1)
    hs = []
    ds = 1
    for lvl in range(len(cfg["channel_mult"])):
        if lvl != 0:
            h = downsample(p, f"{prefix}down.{lvl}.downsample.", h)
            ds *= 2
        for j in range(cfg["num_res_blocks"]):
            h = dilated_resnet_block(p, f"{prefix}down.{lvl}.block.{j}.", h, cfg["num_groups"], (1, 2) if j % 2 == 0 else (4, 8))
            if ds in cfg["attention_resolutions"]:
                h = self_transformer(p, f"{prefix}down.{lvl}.attn.{j}.", h, cfg["num_heads"])
        hs.append(h)
    return hs
