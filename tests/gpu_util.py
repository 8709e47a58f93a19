"""Shared helpers for the -m gpu parity tests."""
import numpy as np

from xtp_b200 import _lib


def ref_contract(desc, A, B, d, C):
    """numpy restatement of the contraction engine's index model (xtp_b200/csrc/contract.cuh)."""
    M, N, K, no, nb = desc.M, desc.N, desc.K, desc.n_outer, desc.n_batch
    b = np.arange(nb)[:, None, None, None]
    o = np.arange(no)[None, :, None, None]
    k = np.arange(K)[None, None, None, :]
    ra = np.arange(M)[None, None, :, None]
    rb = np.arange(N)[None, None, :, None]
    Ai = A[b * desc.a_batch + o * desc.a_outer + ra * desc.a_row + k * desc.a_k]        # (nb,no,M,K)
    Bi = B[b * desc.b_batch + o * desc.b_outer + rb * desc.b_row + k * desc.b_k]        # (nb,no,N,K)
    if d is not None:
        di = d[(b * desc.d_batch + o * desc.d_outer + k)[:, :, 0, :]]                   # (nb,no,K)
        Ai = Ai * di[:, :, None, :]
    prod = np.einsum('bomk,bonk->bmn', Ai, Bi, optimize=True)
    out = C.copy()
    cols = np.arange(N)
    if desc.c_col_inner > 0:
        coff = (cols // desc.c_col_inner) * desc.c_col_outer + (cols % desc.c_col_inner) * desc.c_col
    else:
        coff = cols * desc.c_col
    rows = np.arange(M) * desc.c_row
    for bb in range(nb):
        ix = bb * desc.c_batch + rows[:, None] + coff[None, :]
        new = desc.alpha * prod[bb] + (desc.beta * C[ix] if desc.beta != 0 else 0.0)
        if desc.lower:
            keep = np.arange(M)[:, None] >= cols[None, :]
            out[ix[keep]] = new[keep]
        else:
            out[ix] = new
    return out


def make_desc(**kw):
    d = _lib.ContractDesc()
    d.n_outer = 1
    d.n_batch = 1
    d.alpha = 1.0
    d.beta = 0.0
    d.force_cfg = -1
    d.force_splits = 0
    for k, v in kw.items():
        setattr(d, k, v)
    return d
