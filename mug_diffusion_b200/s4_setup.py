"""Host-side, one-time S4 weight preprocessing: lengthening the persisted C~ of an NPLR kernel.

The reference keeps, per S4 layer, the parameter ``C`` in the transformed form  C~ = C (I - dA^L)  together with an
integer buffer ``L`` (the internal kernel length), and rewrites both IN PLACE the first time a longer sequence is
requested (``SSKernelNPLR._setup_C``, mug/model/s4.py:557-584; called from ``forward`` :726-730):

    L == 0        :  C~ <- C (I - dA^L_new),  L <- L_new                 (fresh, never-run model)
    L_new > L     :  C~ <- C~ (I + dA^L),     L <- 2L   (repeat)          (checkpoint trained at a shorter length)

That is weight preprocessing (parameter-only, once per model state), not per-step work, so it stays on the host like the
reference's own; the per-(model, length) kernel generation itself runs on the GPU (csrc/s4.cu).  dA is the discretised
state matrix built by pushing the 2N unit vectors through the O(N) DPLR step (``_setup_state`` / ``_setup_linear`` /
``_step_state_linear``, s4.py:834-923) and dA^L uses the same square-and-multiply order as ``power`` (s4.py:243-283),
all in complex64 like the reference.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch


def _conj_ext(x: torch.Tensor) -> torch.Tensor:
    """[..., N] -> [..., 2N] by appending the conjugates (s4.py `_conj`)."""
    return torch.cat([x, x.conj()], dim=-1)


def discrete_state_matrix(log_dt: torch.Tensor, inv_w_real: torch.Tensor, w_imag: torch.Tensor, P_ri: torch.Tensor) -> torch.Tensor:
    """dA [H, 2N, 2N] (complex64) of the bilinear-discretised DPLR system A = diag(w) - P P^*, rank 1."""
    dt = torch.exp(log_dt.float())                                            # (H)
    w = (-torch.exp(inv_w_real.float()) + 1j * w_imag.float()).to(torch.complex64)   # (H,N)
    P = torch.view_as_complex(P_ri.float().contiguous())[0]                   # (H,N)   rank-1
    Q = P.conj()
    D = (2.0 / dt[:, None] - w).reciprocal()                                  # (H,N)
    Rs = 1.0 + 2.0 * (Q * D * P).sum(-1).real                                 # (H)     the 1x1 system of _setup_linear
    R = (Q * D) / Rs[:, None]                                                 # (H,N)
    E = 2.0 / dt[:, None] + w                                                 # (H,N)
    Dx, Ex, Px, Qx, Rx = (_conj_ext(t) for t in (D, E, P, Q, R))              # (H,2N)
    n2 = Dx.shape[-1]
    eye = torch.eye(n2, dtype=torch.complex64)                                # rows = unit state vectors e_n
    s = eye[:, None, :]                                                       # (2N, 1, 2N) broadcast over H
    ns = Ex[None] * s - Px[None] * (Qx[None] * s).sum(-1, keepdim=True)       # E s - P (Q . s)
    ns = Dx[None] * (ns - Px[None] * (Rx[None] * ns).sum(-1, keepdim=True))   # D (ns - P (R . ns))
    return ns.permute(1, 2, 0).contiguous()                                   # [n, h, m] -> dA[h, m, n]


def matrix_power_like_reference(A: torch.Tensor, L: int) -> torch.Tensor:
    """A^L with the reference's square-and-multiply order (s4.py `power`, v=None branch)."""
    I = torch.eye(A.shape[-1], dtype=A.dtype).expand_as(A).clone()
    powers = [A]
    while True:
        if L % 2 == 1:
            I = powers[-1] @ I
        L //= 2
        if L == 0:
            break
        powers.append(powers[-1] @ powers[-1])
    return I


def lengthen(params: Dict[str, torch.Tensor], L_internal: int, L_request: int) -> Tuple[torch.Tensor, int]:
    """Return (new C as [1,H,N,2] fp32, new internal length) such that the kernel covers ``L_request``.

    ``params``: 'C' [1,H,N,2], 'log_dt' [H], 'P' [1,H,N,2], 'inv_w_real' [H,N], 'w_imag' [H,N] (CPU tensors)."""
    C = torch.view_as_complex(params["C"].float().contiguous())               # (1,H,N)
    N = C.shape[-1]
    L = int(L_internal)
    if L >= L_request and L > 0:
        return params["C"].clone(), L
    dA = discrete_state_matrix(params["log_dt"], params["inv_w_real"], params["w_imag"], params["P"])
    while L < L_request:
        double = L > 0
        Lp = L if double else int(L_request)
        dA_L = matrix_power_like_reference(dA, Lp)
        Cx = _conj_ext(C)                                                     # (1,H,2N)
        prod = torch.einsum("hnm,chn->chm", dA_L, Cx)                         # row vector times dA^L
        Cx = Cx + prod if double else Cx - prod
        C = Cx[..., :N]
        L = 2 * L if double else Lp
    return torch.view_as_real(C.contiguous()).clone(), L
