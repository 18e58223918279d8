"""Norm fusion of the U-Net plan (round 2): statistics sinks in the tensor-core GEMM epilogue / split-K reduce, single-pass GroupNorm
apply, LayerNorm folded into the Linear behind it -- each against a plain torch fp64 statement of the reference ops
(models.py:10-13 GroupNorm eps 1e-6, attention.py:136-151 LayerNorm eps 1e-5 + Linear), through the C ABI."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from mug_diffusion_b200 import lib as L_  # noqa: E402
from mug_diffusion_b200 import synth  # noqa: E402
from mug_diffusion_b200.engine import OpList  # noqa: E402
from mug_diffusion_b200.packer import _interleave_halves, tf32_split  # noqa: E402

from gpu_util import OpRunner, ncl, nlc, ptr, rel_err, view  # noqa: E402


@pytest.fixture(scope="module")
def R():
    return OpRunner()


def g(name, shape, seed=21):
    return synth._gauss(synth._rng(seed, name), shape)


def _tc_weights(w2d):
    hi, lo = tf32_split(w2d)
    return w2d.cuda(), hi.cuda(), lo.cuda()


@pytest.mark.parametrize("B,L,Cin,Cout,Ctot,col0,split", [(2, 512, 128, 256, 384, 0, 0), (3, 64, 256, 512, 1536, 512, 0), (2, 64, 512, 512, 1536, 1024, 6),
                                                        (5, 48, 128, 128, 256, 128, 0), (3, 124, 256, 384, 384, 0, 3), (2, 256, 64, 64, 128, 64, 0)])
def test_group_moment_sinks(R, B, L, Cin, Cout, Ctot, col0, split):
    """conv3 whose output is columns [col0, col0+Cout) of a Ctot-channel tensor that a GroupNorm(32) will normalise: the epilogue
    (unsplit) or the reduce kernel (split-K) must deliver sum / sum of squares per (sample, group) of exactly what it stored;
    covers groups cut by the window edges, two samples per 128-row tile (L = 48, 64) and ragged tiles (L = 124)"""
    G = 32
    cg = Ctot // G
    x, w, b = g("sx", (B, Cin, L)), g("sw", (Cout, Cin, 3)) / math.sqrt(3 * Cin), 0.1 * g("sb", (Cout,))
    res = g("sr", (B, Cout, L))
    ref = F.conv1d(x.double(), w.double(), b.double(), padding=1) + res.double()
    wp = w.permute(0, 2, 1).contiguous().reshape(Cout, 3 * Cin)
    wc, hc, lc = _tc_weights(wp)
    xc, bc, rc = nlc(x).cuda(), b.cuda(), nlc(res).cuda()
    out = torch.zeros(B * L, Cout).cuda()
    stats = torch.zeros(B, G, 2, dtype=torch.float64).cuda()
    ops = OpList()
    i = ops.gemm(view(xc), ptr(wc), Cout, Cin, view(out), W_hi=ptr(hc), W_lo=ptr(lc), bias=ptr(bc), taps=3, mode=L_.CONV_SAME, Lin=L, Lout=L,
                 residual=view(rc), impl=L_.GEMM_TC, split_k=split)
    assert ops.sink_capable(i)
    ops.add_sink(i, 1, stats.data_ptr(), col0, cg, G)
    R.run(ops)
    assert rel_err(ncl(out.cpu(), B), ref) < 1e-5
    o = ncl(out.cpu(), B).double()                       # moments of what was actually stored
    exp = torch.zeros(B, G, 2, dtype=torch.float64)
    for c in range(Cout):
        gi = (col0 + c) // cg
        exp[:, gi, 0] += o[:, c].sum(-1)
        exp[:, gi, 1] += (o[:, c] ** 2).sum(-1)
    got = stats.cpu()
    assert float((got - exp).abs().max() / exp.abs().max()) < 1e-7       # (a float4's four values are added in fp32 first)


@pytest.mark.parametrize("B,L,K,N,split", [(4, 124, 512, 384, 0), (8, 64, 512, 512, 3), (2, 496, 256, 256, 0)])
def test_group_moment_sinks_on_a_1x1(R, B, L, K, N, split):
    """a Linear / 1x1 conv that feeds a GroupNorm (the transformer block's fused ff_out GEMM): attaching the sink turns it into a
    one-tap conv so that tiles follow the samples; L = 124 is the case where flat 128-row tiles would cut three samples"""
    G = 32
    cg = N // G
    x, w, b = g("px", (B * L, K)), g("pw", (N, K)) / math.sqrt(K), 0.1 * g("pb", (N,))
    ref = F.linear(x.double(), w.double(), b.double())
    wc, hc, lc = _tc_weights(w)
    xc, bc = x.cuda(), b.cuda()
    out = torch.zeros(B * L, N).cuda()
    stats = torch.zeros(B, G, 2, dtype=torch.float64).cuda()
    ops = OpList()
    i = ops.gemm(view(xc), ptr(wc), N, K, view(out), W_hi=ptr(hc), W_lo=ptr(lc), bias=ptr(bc), Lout=L, impl=L_.GEMM_TC, split_k=split)
    assert ops.sink_capable(i) and ops.group_sink_ok(i)
    ops.add_sink(i, 1, stats.data_ptr(), 0, cg, G)
    assert ops.ops[i].u.gemm.conv_mode == L_.CONV_TAPS
    R.run(ops)
    assert rel_err(out, ref) < 1e-5
    o = out.cpu().double().view(B, L, G, cg)
    exp = torch.stack([o.sum((1, 3)), (o ** 2).sum((1, 3))], dim=-1)
    assert float((stats.cpu() - exp).abs().max() / exp.abs().max()) < 1e-7


@pytest.mark.parametrize("M,K,N,split", [(1024, 256, 256, 0), (512, 512, 512, 4), (300, 384, 384, 0), (2048, 128, 64, 0), (640, 256, 512, 2)])
def test_row_moment_sinks(R, M, K, N, split):
    """Linear whose output rows a LayerNorm will normalise: every column tile adds its share of the row's moments"""
    x, w, b = g("rx", (M, K)), g("rw", (N, K)) / math.sqrt(K), 0.1 * g("rb", (N,))
    wc, hc, lc = _tc_weights(w)
    xc, bc = x.cuda(), b.cuda()
    out = torch.zeros(M, N).cuda()
    stats = torch.zeros(M, 2, dtype=torch.float64).cuda()
    ops = OpList()
    i = ops.gemm(view(xc), ptr(wc), N, K, view(out), W_hi=ptr(hc), W_lo=ptr(lc), bias=ptr(bc), impl=L_.GEMM_TC, split_k=split)
    ops.add_sink(i, 2, stats.data_ptr())
    R.run(ops)
    o = out.cpu().double()
    exp = torch.stack([o.sum(-1), (o ** 2).sum(-1)], dim=1)
    assert float((stats.cpu() - exp).abs().max() / exp.abs().max()) < 1e-6        # row sums reduce in fp32 inside a tile


@pytest.mark.parametrize("gate", [L_.GATE_NONE, L_.GATE_GEGLU], ids=["linear", "geglu"])
@pytest.mark.parametrize("M,C,N,split", [(1024, 256, 768, 0), (512, 512, 512, 4), (200, 384, 384, 0), (512, 512, 1024, 0)])
def test_layernorm_folded_into_linear(R, M, C, N, split, gate):
    """Linear(LayerNorm(h)) (attention.py:147-151) as one GEMM on the raw rows: W' = W diag(gamma), epilogue rstd*(acc - mean*colsum) + b',
    with the row moments delivered by the producer of h (here: a first GEMM with a kind-2 sink).  Rows with a large mean included."""
    if gate == L_.GATE_GEGLU and N % 2:
        pytest.skip("gated N must be even")
    a, w0 = g("la", (M, C)), g("lw0", (C, C)) / math.sqrt(C)
    b0 = 3.0 * g("lb0", (C,))                                     # pushes |mean| of h above its spread: the cancellation case
    gam, bet = 1 + 0.2 * g("lg", (C,)), 0.1 * g("lbt", (C,))
    w1, b1 = g("lw1", (N, C)) / math.sqrt(C), 0.1 * g("lb1", (N,))
    h = F.linear(a.double(), w0.double(), b0.double())
    y = F.linear(F.layer_norm(h, (C,), gam.double(), bet.double(), 1e-5), w1.double(), b1.double())
    if gate == L_.GATE_GEGLU:
        v, gt = y.chunk(2, dim=-1)
        y = v * F.gelu(gt)
        w1p, b1p = _interleave_halves(w1), _interleave_halves(b1)
    else:
        w1p, b1p = w1, b1
    wg = (w1p.double() * gam.double()[None]).float()
    colsum = wg.double().sum(1).float()
    bias = (w1p.double() @ bet.double() + b1p.double()).float()
    w0c, h0c, l0c = _tc_weights(w0)
    wgc, hgc, lgc = _tc_weights(wg)
    ac, b0c, csc, bc = a.cuda(), b0.cuda(), colsum.cuda(), bias.cuda()
    hbuf = torch.zeros(M, C).cuda()
    out = torch.zeros(M, N // 2 if gate else N).cuda()
    stats = torch.zeros(M, 2, dtype=torch.float64).cuda()
    ops = OpList()
    i = ops.gemm(view(ac), ptr(w0c), C, C, view(hbuf), W_hi=ptr(h0c), W_lo=ptr(l0c), bias=ptr(b0c), impl=L_.GEMM_TC)
    ops.add_sink(i, 2, stats.data_ptr())
    ops.gemm(view(hbuf), ptr(wgc), N, C, view(out), W_hi=ptr(hgc), W_lo=ptr(lgc), bias=ptr(bc), gate=gate, impl=L_.GEMM_TC, split_k=split,
             ln=(stats.data_ptr(), ptr(csc), 1e-5))
    R.run(ops)
    e = rel_err(out, y)
    print(f"ln-fold M={M} C={C} N={N} split={split} gate={gate} rel_err={e:.2e}")
    assert e < 2e-5


@pytest.mark.parametrize("B,L,C,silu", [(2, 512, 384, True), (3, 64, 1536, True), (2, 124, 512, False), (5, 12, 256, True)])
def test_groupnorm_apply_with_supplied_moments(R, B, L, C, silu):
    """moments accumulated by the stats kernel over two column windows (one cutting a group), then the single-pass apply kernel,
    against torch GroupNorm(32, eps 1e-6) [+ SiLU]"""
    G = 32
    cg = C // G
    x = g("gx", (B, C, L)) * 2 + 0.5
    gam, bet = 1 + 0.1 * g("gg", (C,)), 0.1 * g("gb", (C,))
    ref = F.group_norm(x.double(), G, gam.double(), bet.double(), 1e-6)
    if silu:
        ref = F.silu(ref)
    xc, gc_, bc = nlc(x).cuda(), gam.cuda(), bet.cuda()
    y = torch.zeros(B * L, C).cuda()
    stats = torch.zeros(B, G, 2, dtype=torch.float64).cuda()
    cut = (C // 2 // 4) * 4 + 4 if (C // 2) % cg == 0 else (C // 2 // 4) * 4       # a window edge inside a group when cg > 4
    ops = OpList()
    ops.groupnorm(view(xc, 0, cut), None, 0, 0, B, L, G, False, stats=stats.data_ptr(), stats_col0=0, stats_cg=cg, stats_G=G)
    ops.groupnorm(view(xc, cut, C), None, 0, 0, B, L, G, False, stats=stats.data_ptr(), stats_col0=cut, stats_cg=cg, stats_G=G)
    ops.groupnorm(view(xc), view(y), ptr(gc_), ptr(bc), B, L, G, silu, stats=stats.data_ptr(), stats_col0=0, stats_cg=cg, stats_G=G)
    R.run(ops)
    assert rel_err(ncl(y.cpu(), B), ref) < 1e-5


def test_fused_plan_equals_plain_plan():
    """one U-Net evaluation with the fused norms (default) against the same engine compiled with stand-alone GroupNorm / LayerNorm
    kernels: same network, different launch plan"""
    from mug_diffusion_b200.sampler import MugDiffusionB200
    L, B = 160, 3
    sd = synth.synthetic_state_dict(L)
    inp = synth.synthetic_inputs(B, L, seed=5)
    t = torch.tensor([3, 500, 999]).cuda()
    outs = []
    for fuse in (True, False):
        m = MugDiffusionB200(sd, z_length=L, fuse_norms=fuse)
        outs.append(m.model.forward(inp["x_T"].cuda(), t, inp["c"].cuda(), [w.cuda() for w in inp["w"]]).cpu())
        sess = next(iter(m.engine.sessions.values()))
        kinds = [sess.plan._arr[i].kind for i in range(sess.plan.n_ops)]
        assert (kinds.count(L_.OP_LAYERNORM) == 0) == fuse
        del m
    assert rel_err(outs[0], outs[1]) < 2e-5
