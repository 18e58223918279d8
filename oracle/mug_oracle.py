"""CPU oracle for the Mug-Diffusion denoising hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this
file.  The product path (mug_diffusion_b200/) never does; it fails loudly when the CUDA library is absent.

What it is: a functional fp32 restatement (plain torch CPU ops over a flat ``{name: tensor}`` state dict
that uses the reference's own state_dict key names) of the one path BASELINE.json names:

    DDIMSampler.sample -> p_sample_ddim -> UNetModel.forward -> ... -> AutoencoderKL.decode

All arithmetic is floating point, so following the tier rules the oracle is a torch-fp32 restatement; the
third-party arithmetic it leans on is ATen CPU (torch 2.11.0, pinned by the image) -- the same library
"the reference PyTorch path" means.

Parity pin: the reference ships no tests or golden vectors, so the pins were created by
running the UNMODIFIED reference in the build container (tools/ref_shim.py + tools/make_goldens.py) on
the seeded synthetic weights of mug_diffusion_b200/synth.py and committing its outputs under
tests/golden/.  tests/test_oracle_golden.py checks this file against those vectors.

Every function cites the reference file:line (paths relative to /root/reference) it follows.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

# --------------------------------------------------------------------------------------------------
# default architecture = configs/mug/mug_diffusion.yaml:28-58
# --------------------------------------------------------------------------------------------------
DEFAULT_UNET = dict(
    in_channels=16, model_channels=128, out_channels=16, attention_resolutions=(8, 4, 2),
    num_res_blocks=2, channel_mult=(1, 2, 3, 4), num_heads=8, context_dim=128,
    audio_channels=(256, 512, 512, 512), s4_layer=True, s4_state=64, pos_max=64,
)
DEFAULT_DECODER = dict(
    x_channels=16, middle_channels=64, z_channels=16, num_groups=8, channel_mult=(1, 2, 4, 4),
    num_res_blocks=1, scale=1.0,
)
UNET_PREFIX = "model.unet_model."
DECODER_PREFIX = "model.first_stage_model.decoder."


# --------------------------------------------------------------------------------------------------
# leaf ops
# --------------------------------------------------------------------------------------------------
def group_norm(p: Params, pre: str, x: torch.Tensor, groups: int) -> torch.Tensor:
    """Normalize() = GroupNorm(groups, C, eps=1e-6, affine)  -- mug/model/models.py:10-13."""
    return F.group_norm(x, groups, p[pre + "weight"], p[pre + "bias"], eps=1e-6)


def conv1d(p: Params, pre: str, x: torch.Tensor, stride: int = 1, padding: int = 0) -> torch.Tensor:
    return F.conv1d(x, p[pre + "weight"], p.get(pre + "bias"), stride=stride, padding=padding)


def linear(p: Params, pre: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, p[pre + "weight"], p.get(pre + "bias"))


def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    """[cos | sin] sinusoid table  -- mug/model/util.py:156-176."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def downsample(p: Params, pre: str, x: torch.Tensor) -> torch.Tensor:
    """right-pad 1, conv3 stride 2  -- mug/model/models.py:84-91."""
    return conv1d(p, pre + "conv.", F.pad(x, (0, 1)), stride=2)


def upsample(p: Params, pre: str, x: torch.Tensor) -> torch.Tensor:
    """nearest x2 then conv3 pad 1  -- mug/model/models.py:66-70."""
    x = x.repeat_interleave(2, dim=-1)
    return conv1d(p, pre + "conv.", x, padding=1)


# --------------------------------------------------------------------------------------------------
# TimestepResBlock / ResnetBlock
# --------------------------------------------------------------------------------------------------
def timestep_resblock(p: Params, pre: str, x: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    """GN32-SiLU-conv3, + Linear(SiLU(emb)), GN32-SiLU-conv3, + skip  -- mug/diffusion/unet.py:212-239."""
    h = conv1d(p, pre + "in_layers.2.", F.silu(group_norm(p, pre + "in_layers.0.", x, 32)), padding=1)
    h = h + linear(p, pre + "emb_layers.1.", F.silu(emb))[:, :, None]
    h = conv1d(p, pre + "out_layers.3.", F.silu(group_norm(p, pre + "out_layers.0.", h, 32)), padding=1)
    if pre + "skip_connection.weight" in p:
        x = conv1d(p, pre + "skip_connection.", x)
    return x + h


def resnet_block(p: Params, pre: str, x: torch.Tensor, groups: int) -> torch.Tensor:
    """Decoder ResnetBlock without time embedding  -- mug/model/models.py:142-159."""
    h = conv1d(p, pre + "conv1.", F.silu(group_norm(p, pre + "norm1.", x, groups)), padding=1)
    h = conv1d(p, pre + "conv2.", F.silu(group_norm(p, pre + "norm2.", h, groups)), padding=1)
    if pre + "nin_shortcut.weight" in p:
        x = conv1d(p, pre + "nin_shortcut.", x)
    return x + h


# --------------------------------------------------------------------------------------------------
# attention
# --------------------------------------------------------------------------------------------------
def cross_attention(p: Params, pre: str, x: torch.Tensor, context: Optional[torch.Tensor],
                    heads: int, pos_max: int = 64) -> torch.Tensor:
    """CrossAttention.forward  -- mug/model/attention.py:91-126.

    sim = (q k^T + relpos[idx]) * d^-1/2 ;  attn = softmax(sim) * C_embedding[idx] ;  out = attn v.
    idx = clamp(j - i, -pos_max, pos_max) + pos_max   (j = key index, i = query index)
    x: [B, Lq, C]; context: [B, Lk, Cc] or None (self attention).
    """
    ctx = x if context is None else context
    q = F.linear(x, p[pre + "to_q.weight"])
    k = F.linear(ctx, p[pre + "to_k.weight"])
    v = F.linear(ctx, p[pre + "to_v.weight"])
    B, Lq, inner = q.shape
    Lk = k.shape[1]
    d = inner // heads
    q = q.view(B, Lq, heads, d).permute(0, 2, 1, 3)
    k = k.view(B, Lk, heads, d).permute(0, 2, 1, 3)
    v = v.view(B, Lk, heads, d).permute(0, 2, 1, 3)
    idx = (torch.arange(Lk)[None, :] - torch.arange(Lq)[:, None]).clamp(-pos_max, pos_max) + pos_max
    bias = p[pre + "relative_position_embedding"][idx].permute(2, 0, 1)  # [h, Lq, Lk]
    cmul = p[pre + "C_embedding"][idx].permute(2, 0, 1)
    sim = (torch.matmul(q, k.transpose(-1, -2)) + bias[None]) * (d ** -0.5)
    attn = sim.softmax(dim=-1) * cmul[None]
    out = torch.matmul(attn, v).permute(0, 2, 1, 3).reshape(B, Lq, inner)
    return linear(p, pre + "to_out.0.", out)


def transformer_block(p: Params, pre: str, x: torch.Tensor, context: torch.Tensor, heads: int) -> torch.Tensor:
    """LN-selfattn-+, LN-crossattn-+, LN-GEGLU FF-+  -- mug/model/attention.py:147-151, 38-65."""
    C = x.shape[-1]

    def ln(name, y):
        return F.layer_norm(y, (C,), p[pre + name + ".weight"], p[pre + name + ".bias"], eps=1e-5)

    x = cross_attention(p, pre + "attn1.", ln("norm1", x), None, heads) + x
    x = cross_attention(p, pre + "attn2.", ln("norm2", x), context, heads) + x
    proj = linear(p, pre + "ff.net.0.proj.", ln("norm3", x))
    a, gate = proj.chunk(2, dim=-1)
    x = linear(p, pre + "ff.net.2.", a * F.gelu(gate)) + x
    return x


def contextual_transformer(p: Params, pre: str, x: torch.Tensor, context: torch.Tensor, heads: int) -> torch.Tensor:
    """GN32 -> 1x1 conv -> [B,L,C] -> block -> 1x1 conv -> +x  -- mug/model/attention.py:186-199.
    x: [B,C,L]; context: [B,Cc,Lc] (channel-first, as the reference passes it)."""
    h = conv1d(p, pre + "proj_in.", group_norm(p, pre + "norm.", x, 32))
    h = transformer_block(p, pre + "transformer_blocks.0.", h.transpose(1, 2), context.transpose(1, 2), heads)
    return conv1d(p, pre + "proj_out.", h.transpose(1, 2)) + x


# --------------------------------------------------------------------------------------------------
# S4
# --------------------------------------------------------------------------------------------------
def s4_nplr_kernel(p: Params, pre: str, L: int) -> torch.Tensor:
    """SSKernelNPLR.forward for rank 1, channels 1, rate 1, state None, already-initialised internal
    length (buffer ``L`` >= requested L; no _setup_C call)  -- mug/model/s4.py:706-832 with the
    non-conjugate Cauchy sum ``cauchy_naive`` (s4.py:140-147) the reference falls back to.

    ``pre`` ends in 'kernel.kernel.'.  Returns K [H, L] float32.
    """
    L_int = int(p[pre + "L"].item())
    if L > L_int:
        raise ValueError(f"S4 kernel length {L} > stored internal length {L_int}: run s4_setup_C first")
    dt = torch.exp(p[pre + "log_dt"])  # (H)
    Bc = torch.view_as_complex(p[pre + "B"])  # (1,H,N)
    Cc = torch.view_as_complex(p[pre + "C"])  # (1,H,N)
    Pc = torch.view_as_complex(p[pre + "P"])  # (1,H,N)
    Qc = Pc.conj()
    w = (-torch.exp(p[pre + "inv_w_real"]) + 1j * p[pre + "w_imag"]).to(torch.complex64)  # s4.py:689-704
    # FFT nodes (s4.py:586-604): complex64 power of a complex64 base, as the reference computes them
    omega = torch.tensor(np.exp(-2j * np.pi / L_int), dtype=torch.complex64) ** torch.arange(0, L_int // 2 + 1)
    z = 2 * (1 - omega) / (1 + omega)
    w = w * dt[:, None]
    Bs = torch.cat([Bc, Pc], dim=0)  # (2,H,N)
    Cs = torch.cat([Cc, Qc], dim=0)  # (2,H,N)
    v = Bs[:, None] * Cs[None, :]  # (2,2,H,N)
    r = (v[..., :, None] / (z[None, None, None, None, :] - w[None, None, :, :, None])).sum(dim=-2)  # (2,2,H,Lf)
    r = r * dt[None, None, :, None]
    k_f = r[:-1, :-1] - r[:-1, -1:] * r[-1:, :-1] / (1 + r[-1:, -1:])  # Woodbury, s4.py:791-792
    k_f = k_f * 2 / (1 + omega)
    k = torch.fft.irfft(k_f, n=L_int)[..., :L]
    return k[0, 0]


def s4_block(p: Params, pre: str, u: torch.Tensor) -> torch.Tensor:
    """S4.forward (transposed, unidirectional, gelu + conv1x1 + GLU)  -- mug/model/s4.py:1471-1541.
    ``pre`` ends in 's4_model.'.  u: [B,H,L]."""
    L = u.shape[-1]
    k = s4_nplr_kernel(p, pre + "kernel.kernel.", L)  # [H,L]
    k_f = torch.fft.rfft(k, n=2 * L)
    u_f = torch.fft.rfft(u, n=2 * L)
    y = torch.fft.irfft(u_f * k_f[None], n=2 * L)[..., :L]
    y = y + u * p[pre + "D"][0][None, :, None]
    y = F.gelu(y)
    y = conv1d(p, pre + "output_linear.0.", y)
    a, b = y.chunk(2, dim=1)  # nn.GLU(dim=-2)
    return a * torch.sigmoid(b)


def s4_layer(p: Params, pre: str, x: torch.Tensor) -> torch.Tensor:
    """GN32 -> S4 -> conv3 -> +x  -- mug/diffusion/unet.py:86-91."""
    h = s4_block(p, pre + "s4_model.", group_norm(p, pre + "norm.", x, 32))
    return x + conv1d(p, pre + "out_layer.", h, padding=1)


# --------------------------------------------------------------------------------------------------
# U-Net
# --------------------------------------------------------------------------------------------------
def unet_layout(cfg: dict) -> dict:
    """Walk the constructor (mug/diffusion/unet.py:341-493) and return the block structure:
    lists of (kind, ...) tuples for input_blocks / middle / output_blocks."""
    mc, mult, nrb = cfg["model_channels"], cfg["channel_mult"], cfg["num_res_blocks"]
    att, heads, s4 = set(cfg["attention_resolutions"]), cfg["num_heads"], cfg["s4_layer"]
    inp: List[list] = [[("conv_in",)]]
    ds = 1
    for level, m in enumerate(mult):
        inp.append([("audio",)])
        for _ in range(nrb):
            layers = [("res",)]
            if ds in att:
                layers.append(("attn", heads))
            if s4:
                layers.append(("s4",))
            inp.append(layers)
        if level != len(mult) - 1:
            inp.append([("down",)])
            ds *= 2
    mid = [("res",), ("attn", heads), ("res",)]
    out: List[list] = []
    for level in reversed(range(len(mult))):
        out.append([("audio",)])
        for i in range(nrb + 1):
            layers = [("res",)]
            if ds in att:
                layers.append(("attn", heads))
            if s4 and i != nrb:
                layers.append(("s4",))
            if level and i == nrb:
                layers.append(("up",))
                ds //= 2
            out.append(layers)
    return dict(input=inp, middle=mid, output=out)


def _run_layers(p, pre, layers, h, emb, context):
    for j, layer in enumerate(layers):
        lp = f"{pre}{j}."
        kind = layer[0]
        if kind == "conv_in":
            h = conv1d(p, lp, h, padding=1)
        elif kind == "res":
            h = timestep_resblock(p, lp, h, emb)
        elif kind == "attn":
            h = contextual_transformer(p, lp, h, context, layer[1])
        elif kind == "s4":
            h = s4_layer(p, lp, h)
        elif kind == "down":
            h = downsample(p, lp, h)
        elif kind == "up":
            h = upsample(p, lp, h)
        else:
            raise ValueError(kind)
    return h


def unet_forward(p: Params, x: torch.Tensor, t: torch.Tensor, context: torch.Tensor,
                 audios: Sequence[torch.Tensor], cfg: dict = DEFAULT_UNET, prefix: str = UNET_PREFIX) -> torch.Tensor:
    """UNetModel.forward  -- mug/diffusion/unet.py:511-550.
    x [B,16,L]; t [B] long; context [B,128,21]; audios: the wave-encoder output list (last 4 used)."""
    lay = unet_layout(cfg)
    nlev = len(cfg["channel_mult"])
    emb = timestep_embedding(t, cfg["model_channels"])
    emb = linear(p, prefix + "time_embed.2.", F.silu(linear(p, prefix + "time_embed.0.", emb)))
    hs = []
    h = x
    ai = -nlev
    for i, layers in enumerate(lay["input"]):
        if layers[0][0] == "audio":
            h = torch.cat([h, audios[ai]], dim=1)
            ai += 1
        else:
            h = _run_layers(p, f"{prefix}input_blocks.{i}.", layers, h, emb, context)
            hs.append(h)
    ai = -1
    h = _run_layers(p, prefix + "middle_block.", lay["middle"], h, emb, context)
    for i, layers in enumerate(lay["output"]):
        if layers[0][0] == "audio":
            h = torch.cat([h, audios[ai]], dim=1)
            ai -= 1
        else:
            h = torch.cat([h, hs.pop()], dim=1)
            h = _run_layers(p, f"{prefix}output_blocks.{i}.", layers, h, emb, context)
    h = F.silu(group_norm(p, prefix + "out.0.", h, 32))
    return conv1d(p, prefix + "out.2.", h, padding=1)


# --------------------------------------------------------------------------------------------------
# first-stage decoder
# --------------------------------------------------------------------------------------------------
def decoder_forward(p: Params, z: torch.Tensor, cfg: dict = DEFAULT_DECODER, prefix: str = DECODER_PREFIX) -> torch.Tensor:
    """AutoencoderKL.decode + Decoder.forward  -- mug/firststage/autoencoder.py:75-77, 329-354."""
    g = cfg["num_groups"]
    nres = len(cfg["channel_mult"])
    h = conv1d(p, prefix + "conv_in.", z / cfg.get("scale", 1.0), padding=1)
    h = resnet_block(p, prefix + "mid.block_1.", h, g)
    h = resnet_block(p, prefix + "mid.block_2.", h, g)
    for lvl in reversed(range(nres)):
        for b in range(cfg["num_res_blocks"] + 1):
            h = resnet_block(p, f"{prefix}up.{lvl}.block.{b}.", h, g)
        if lvl != 0:
            h = upsample(p, f"{prefix}up.{lvl}.upsample.", h)
    h = F.silu(group_norm(p, prefix + "norm_out.", h, g))
    return conv1d(p, prefix + "conv_out.", h, padding=1)


# --------------------------------------------------------------------------------------------------
# schedule + DDIM sampler
# --------------------------------------------------------------------------------------------------
def make_schedule(S: int, eta: float = 0.0, timesteps: int = 1000, linear_start: float = 1e-4,
                  linear_end: float = 2e-2) -> dict:
    """beta schedule + DDIM tables, same mixed f64/f32 arithmetic as the reference:
    mug/diffusion/utils.py:16-40 (betas f64), diffusion.py:131-151 (alphas_cumprod -> f32),
    ddim.py:24-53 + utils.py:50-80 (uniform timesteps, alphas, alphas_prev, sigmas)."""
    betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=torch.float64) ** 2).numpy()
    acp64 = np.cumprod(1.0 - betas, axis=0)
    alphas_cumprod = torch.tensor(acp64, dtype=torch.float32)
    c = timesteps // S
    ddim_timesteps = np.asarray(list(range(0, timesteps, c))) + 1
    ac = alphas_cumprod  # float32 torch tensor indexed with numpy ints, as ddim.py does
    alphas = ac[ddim_timesteps]
    alphas_prev = np.asarray([ac[0]] + ac[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return dict(timesteps=ddim_timesteps, alphas=alphas, alphas_prev=alphas_prev, sigmas=sigmas,
                sqrt_one_minus_alphas=np.sqrt(1.0 - alphas), alphas_cumprod=alphas_cumprod,
                # q_sample tables: square roots taken in f64, then cast (diffusion.py:148-150)
                sqrt_alphas_cumprod=torch.tensor(np.sqrt(acp64), dtype=torch.float32),
                sqrt_one_minus_alphas_cumprod=torch.tensor(np.sqrt(1.0 - acp64), dtype=torch.float32))


def ddim_sample(p: Params, S: int, c: torch.Tensor, w: Sequence[torch.Tensor], x_T: torch.Tensor,
                scale: float = 1.0, uc: Optional[torch.Tensor] = None, eta: float = 0.0,
                cfg: dict = DEFAULT_UNET, noise_gen: Optional[torch.Generator] = None,
                return_eps: bool = False, noise_seq: Optional[Sequence[torch.Tensor]] = None, temperature: float = 1.0,
                mask: Optional[torch.Tensor] = None, x0: Optional[torch.Tensor] = None,
                q_noise_seq: Optional[Sequence[torch.Tensor]] = None):
    """DDIMSampler.ddim_sampling + p_sample_ddim  -- mug/diffusion/ddim.py:110-196.
    ``noise_seq[i]`` replaces the i-th ``noise_like`` draw (:192), already passed through ``dropout`` (:193-194) if any, so a
    test can hand the exact per-step noise of another RNG stream to this restatement.
    ``mask`` / ``x0``: the inpainting blend of :140-143 with ``DDPM.q_sample`` (diffusion.py:327-333); ``q_noise_seq[i]`` replaces the
    ``randn_like(x0)`` it draws in iteration i."""
    sch = make_schedule(S, eta)
    ts = sch["timesteps"]
    x = x_T
    B = x.shape[0]
    total = ts.shape[0]
    eps_list = []
    for i, step in enumerate(np.flip(ts)):
        index = total - i - 1
        t = torch.full((B,), int(step), dtype=torch.long)
        if mask is not None:
            assert x0 is not None
            qn = q_noise_seq[i] if q_noise_seq is not None else torch.randn(x0.shape, generator=noise_gen)
            x_orig = sch["sqrt_alphas_cumprod"][t].view(-1, 1, 1) * x0 + sch["sqrt_one_minus_alphas_cumprod"][t].view(-1, 1, 1) * qn
            x = x_orig * mask + (1. - mask) * x
        if uc is None or scale == 1.0:
            e_t = unet_forward(p, x, t, c, w, cfg)
        else:
            e = unet_forward(p, torch.cat([x, x]), torch.cat([t, t]), torch.cat([uc, c]),
                             [torch.cat([wi, wi]) for wi in w], cfg)
            e_u, e_c = e.chunk(2)
            e_t = e_u + scale * (e_c - e_u)
        if return_eps:
            eps_list.append(e_t)
        a_t = torch.full((B, 1, 1), float(sch["alphas"][index]))
        a_prev = torch.full((B, 1, 1), float(sch["alphas_prev"][index]))
        sigma_t = torch.full((B, 1, 1), float(sch["sigmas"][index]))
        s1m = torch.full((B, 1, 1), float(sch["sqrt_one_minus_alphas"][index]))
        pred_x0 = (x - s1m * e_t) / a_t.sqrt()
        dir_xt = (1.0 - a_prev - sigma_t ** 2).sqrt() * e_t
        raw = noise_seq[i] if noise_seq is not None else torch.randn(x.shape, generator=noise_gen)
        noise = sigma_t * raw * temperature                                  # :192
        x = a_prev.sqrt() * pred_x0 + dir_xt + noise
    if return_eps:
        return x, eps_list
    return x


def notes_from_logits(logits: torch.Tensor) -> torch.Tensor:
    """The note on/off decisions the .osu writer takes from decoder logits [B,16,T] for 4K:
    OsuManiaConvertor.array_to_objects thresholds rows 0-3 (is_start) and 8-11 (is_holding) with
    ``> 0`` (from_logits) -- mug/data/convertor.py:211-216, 232-264.  Returns bool [B,8,T]."""
    return torch.cat([logits[:, 0:4], logits[:, 8:12]], dim=1) > 0


def array_to_objects(note_array: np.ndarray, key_count: int, frame_ms: float) -> List[str]:
    """OsuManiaConvertor.array_to_objects with from_logits=True  -- mug/data/convertor.py:211-264.
    note_array: [4*key_count, T] decoder logits of one chart.  Returns the .osu hit-object lines sorted by start time
    (stable), exactly as the reference writes them."""
    a = np.asarray(note_array).transpose()                      # [T, 4K]
    T = a.shape[0]
    K = int(key_count)
    width = int(512 / K)
    out = []
    for col in range(K):
        for si in np.where(a[:, col] > 0)[0]:
            so = np.clip(a[si, col + K], 0, 1)
            start = int(round((si + so) * frame_ms))
            end = -1
            if si != T - 1:
                i = si + 1
                while i < T and a[i, col + 2 * K] > 0 and not a[i, col] > 0:
                    i += 1
                ei = i - 1
                if ei != si:
                    eo = np.clip(a[ei, col + 3 * K], 0, 1)
                    end = int(round((ei + eo) * frame_ms))
            x = int(round((col + 0.5) * width))
            line = f"{x},192,{start},1,0,0:0:0:0:" if end == -1 else f"{x},192,{start},128,0,{end}:0:0:0:0:"
            out.append((line, start))
    out.sort(key=lambda t: t[1])
    return [t[0] for t in out]


# ---------------------------------------------------------------------------------------------------
# Prompt path.  TEST INFRASTRUCTURE like the rest of this file.
# ---------------------------------------------------------------------------------------------------
def feature_rows(x):
    """mug/util.py:50-60 count_beatmap_features_embedding"""
    import math
    if x["type"] == "numeric":
        return int(math.ceil((x["max"] - x["min"]) / x["interval"])) + 1
    if x["type"] == "category":
        return len(x["category"]) + 1
    if x["type"] == "bool":
        return 3
    raise ValueError(str(x))


def feature_ids(feature_dict, feature_yaml):
    """mug/util.py:62-84 feature_dict_to_embedding_ids, restated as (offset of the slot's row block) + (bin within it)"""
    out, offset = [], 0
    for spec in feature_yaml:
        v = feature_dict.get(spec["name"])
        if v is None:
            b = 0                                                       # :67-68 missing
        elif spec["type"] == "numeric":
            v = max(spec["min"], min(spec["max"], v))                   # :71
            b = int((v - spec["min"]) / spec["interval"]) + 1           # :72, :80
        elif spec["type"] == "bool":
            b = v + 1                                                   # :74, :80
        else:
            b = spec["category"].index(v) + 1                           # :77 (ValueError for an unknown value), :80
        for _ in range(spec.get("count", 1)):                           # :81-83
            out.append(offset + b)
            offset += feature_rows(spec)
    return out


def prompt_embed(table: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """mug/cond/feature.py:15-21: embedding(x.long()) then 'b f h -> b h f'"""
    return table[ids.long()].permute(0, 2, 1).contiguous()
