"""numpy mirror of the compressed Sigma_c grid scan (xtp_b200/csrc/kernels.cu section (1b)): the same bin table,
Chebyshev moments, near/far split and series as the CUDA kernels, statement by statement, fed by the REAL host plan
(xtpb_ppm_grid_plan).  Test infrastructure only: it lets the CPU suite check the plan logic and the mathematics of the
kernels against the oracle's direct pole sum without a GPU."""
import ctypes as C

import numpy as np

from oracle import gwbse_oracle as orc
from xtp_b200 import _lib

ORDER = 16       # kCmpOrder
CHUNK = _lib.lib().xtpb_ppm_grid_chunk()       # kCmpChunk


def plan(grid_start, spacing, steps, zmin, zmax):
    lib = _lib.lib()
    grid_start = np.ascontiguousarray(grid_start, dtype=np.float64)
    nl = len(grid_start)
    nb, nch, usable = _lib.idx(0), _lib.idx(0), C.c_int(0)
    dp = lambda a: a.ctypes.data_as(_lib.dptr)
    _lib.check(lib.xtpb_ppm_grid_plan(nl, dp(grid_start), float(spacing), int(steps), float(zmin), float(zmax), 0, None,
                                      C.byref(nb), None, C.byref(nch), C.byref(usable)))
    if not usable.value:
        return None
    edges = np.empty(nb.value + 1)
    near = np.empty((nl, nch.value, 4), dtype=np.int32)
    _lib.check(lib.xtpb_ppm_grid_plan(nl, dp(grid_start), float(spacing), int(steps), float(zmin), float(zmax),
                                      len(edges), dp(edges), C.byref(nb), near.ctypes.data_as(C.POINTER(C.c_int)),
                                      C.byref(nch), C.byref(usable)))
    return edges, near


def pole_range(e, n_occ, freq, fac):
    live = fac != 0.0
    lo, hi = freq[live].min(), freq[live].max()
    zmin, zmax = np.inf, -np.inf
    if n_occ > 0:
        zmin, zmax = min(zmin, e[0] - hi), max(zmax, e[n_occ - 1] - lo)
    if n_occ < len(e):
        zmin, zmax = min(zmin, e[n_occ] + lo), max(zmax, e[-1] + hi)
    return zmin, zmax


def bin_table(edges, e, n_occ, freq):
    """ppm_bin_table_kernel: table[seg, P, b] = first m of the segment with e[m] + shift >= edges[b]."""
    nb, naux, nt = len(edges) - 1, len(freq), len(e)
    table = np.empty((2, naux, nb + 1), dtype=np.int64)
    for seg, (lo, hi) in enumerate(((0, n_occ), (n_occ, nt))):
        for P in range(naux):
            z = e[lo:hi] + (freq[P] if seg else -freq[P])
            table[seg, P, 1:nb] = lo + np.searchsorted(z, edges[1:nb], side="left")
            table[seg, P, 0], table[seg, P, nb] = lo, hi
    return table


def moments(slab, e, freq, fac, edges, table):
    """ppm_moments_kernel: mu[b, j] = sum over the poles of bin b of fac_P M^2 T_j((z - c) / h)."""
    nb = len(edges) - 1
    mu = np.zeros((nb, ORDER))
    for b in range(nb):
        c, hinv = 0.5 * (edges[b] + edges[b + 1]), 2.0 / (edges[b + 1] - edges[b])
        for P in np.nonzero(fac)[0]:
            for seg in range(2):
                lo, hi = table[seg, P, b], table[seg, P, b + 1]
                if hi <= lo:
                    continue
                a = fac[P] * slab[P, lo:hi] ** 2
                t = (e[lo:hi] + (freq[P] if seg else -freq[P]) - c) * hinv
                tm, tc = np.ones_like(t), t
                mu[b, 0] += a.sum()
                mu[b, 1] += (a * t).sum()
                for j in range(2, ORDER):
                    tm, tc = tc, 2.0 * t * tc - tm
                    mu[b, j] += (a * tc).sum()
    return mu


def equivalent_poles(mu, edges):
    """ppm_equivalent_poles_kernel: per bin, ORDER poles at the Chebyshev-Gauss nodes with weights
    w_n = (2/K) [mu_0/2 + sum_k mu_k T_k(t_n)] that reproduce sum_i A_i f(z_i) for every f smooth on the bin."""
    nb = len(edges) - 1
    n = np.arange(ORDER)
    t = np.cos(np.pi * (n + 0.5) / ORDER)
    Tk = np.cos(np.outer(np.arange(ORDER), np.arccos(t)))          # T_k(t_n)
    w = (2.0 / ORDER) * (0.5 * mu[:, :1] + mu[:, 1:] @ Tk[1:])        # (nb, ORDER)
    c, h = 0.5 * (edges[:-1] + edges[1:]), 0.5 * np.diff(edges)
    return w, c[:, None] + h[:, None] * t[None, :]


def grid_values(slab, e, n_occ, freq, fac, om0, spacing, steps, edges, near_level, counters=None):
    """sigma_ppm_grid_compressed_kernel for one level: poles of the near bins one by one (damped kernel) except the
    INNER bins (inside every target's damping window), which enter through their equivalent poles; far bins by the
    Cauchy series of their moments."""
    table = bin_table(edges, e, n_occ, freq)
    mu = moments(slab, e, freq, fac, edges, table)
    eq_w, eq_z = equivalent_poles(mu, edges)
    nb = len(edges) - 1
    out = np.zeros(steps)
    for ch in range(near_level.shape[0]):
        js = np.arange(ch * CHUNK, min(steps, (ch + 1) * CHUNK))
        om = om0 + spacing * js
        b_lo, b_hi, i_lo, i_hi = near_level[ch]
        acc = np.zeros(len(js))
        if b_lo <= b_hi:
            for P in np.nonzero(fac)[0]:
                for seg in range(2):
                    for lo, hi in ((table[seg, P, b_lo], table[seg, P, i_lo]), (table[seg, P, i_hi + 1], table[seg, P, b_hi + 1])):
                        if hi <= lo:
                            continue
                        z = e[lo:hi] + (freq[P] if seg else -freq[P])
                        a = fac[P] * slab[P, lo:hi] ** 2
                        acc += (a[None, :] * orc.ppm_stabilized_inverse(om[:, None] - z[None, :])).sum(axis=1)
                        if counters is not None:
                            counters["near"] += (hi - lo) * len(js)
            for b in range(i_lo, i_hi + 1):
                x = om[:, None] - eq_z[b][None, :]
                assert np.all(np.abs(x) < 0.25), "an inner bin must lie inside every target's damping window"
                acc += (eq_w[b][None, :] * orc.ppm_stabilized_inverse(x)).sum(axis=1)
        for b in range(nb):
            if b_lo <= b <= b_hi:
                continue
            h = 0.5 * (edges[b + 1] - edges[b])
            d = (om - 0.5 * (edges[b] + edges[b + 1])) / h
            ad, sg = np.abs(d), np.where(d < 0.0, -1.0, 1.0)
            sq = np.sqrt(ad * ad - 1.0)
            r = sg / (ad + sq)
            f = np.zeros(len(js))
            for j in range(ORDER - 1, 0, -1):
                f = (f + mu[b, j]) * r
            f = 0.5 * mu[b, 0] + f
            acc += f * 2.0 * sg / (sq * h)
        out[js] = acc
    if counters is not None:
        counters["all"] += int(np.count_nonzero(fac)) * len(e) * steps
    return out


def direct_values(slab, e, n_occ, freq, fac, om0, spacing, steps):
    """the oracle's pole-by-pole sum (Sigma_PPM::CalcCorrelationDiagElement) on the same grid"""
    z = np.where(np.arange(len(e))[None, :] < n_occ, e[None, :] - freq[:, None], e[None, :] + freq[:, None])
    a = fac[:, None] * slab * slab
    return np.array([(a * orc.ppm_stabilized_inverse(om0 + spacing * j - z)).sum() for j in range(steps)])
