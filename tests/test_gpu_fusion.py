"""LayerNorm fold of the U-Net plan (round 2): the producer of a LayerNorm's input accumulates the row moments in its epilogue /
split-K reduce, the Linear behind the LayerNorm runs on the raw rows and corrects in its epilogue -- against a plain torch fp64
statement of the reference ops (attention.py:136-151 LayerNorm eps 1e-5 + Linear), through the C ABI."""
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
    assert ops.can_deliver_row_moments(i)
    ops.ops[i].u.gemm.row_moments = stats.data_ptr()
    R.run(ops)
    o = out.cpu().double()
    exp = torch.stack([o.sum(-1), (o ** 2).sum(-1)], dim=1)
    assert float((stats.cpu() - exp).abs().max() / exp.abs().max()) < 1e-6        # row sums reduce in fp32 inside a tile


@pytest.mark.parametrize("gate", [L_.GATE_NONE, L_.GATE_GEGLU], ids=["linear", "geglu"])
@pytest.mark.parametrize("M,C,N,split", [(1024, 256, 768, 0), (512, 512, 512, 4), (200, 384, 384, 0), (512, 512, 1024, 0)])
def test_layernorm_folded_into_linear(R, M, C, N, split, gate):
    """Linear(LayerNorm(h)) (attention.py:147-151) as one GEMM on the raw rows: W' = W diag(gamma), epilogue rstd*(acc - mean*colsum) + b',
    with the row moments delivered by the producer of h (here: a first GEMM with row_moments set).  Rows with a large mean included."""
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
    ops.ops[i].u.gemm.row_moments = stats.data_ptr()
    ops.gemm(view(hbuf), ptr(wgc), N, C, view(out), W_hi=ptr(hgc), W_lo=ptr(lgc), bias=ptr(bc), gate=gate, impl=L_.GEMM_TC, split_k=split,
             ln=(stats.data_ptr(), ptr(csc), 1e-5))
    R.run(ops)
    e = rel_err(out, y)
    print(f"ln-fold M={M} C={C} N={N} split={split} gate={gate} rel_err={e:.2e}")
    assert e < 2e-5


def test_folded_plan_equals_plain_plan():
    """one U-Net evaluation with every LayerNorm folded against the same engine compiled with stand-alone LayerNorm kernels:
    same network, different launch plan"""
    from mug_diffusion_b200.sampler import MugDiffusionB200
    L, B = 160, 3
    sd = synth.synthetic_state_dict(L)
    inp = synth.synthetic_inputs(B, L, seed=5)
    t = torch.tensor([3, 500, 999]).cuda()
    outs = []
    for fuse in (True, False):
        m = MugDiffusionB200(sd, z_length=L, fold_ln=fuse)
        outs.append(m.model.forward(inp["x_T"].cuda(), t, inp["c"].cuda(), [w.cuda() for w in inp["w"]]).cpu())
        sess = next(iter(m.engine.sessions.values()))
        kinds = [sess.plan._arr[i].kind for i in range(sess.plan.n_ops)]
        assert (kinds.count(L_.OP_LAYERNORM) == 0) == fuse
        del m
    assert rel_err(outs[0], outs[1]) < 2e-5
