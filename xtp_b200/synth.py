"""Synthetic GW-BSE inputs (SURVEY.md section 8d recipe, BASELINE.json configs[4]).

Real-molecule inputs (AO three-centre integrals, DFT orbitals, Vxc) come from
XTP's integral/DFT layers, which are outside the hot path; these generators
produce inputs of the same shapes and magnitudes so that every stage of the
path (M build, RPA, Sigma, BSE) runs on well-conditioned numbers:
  * MO coefficients: orthonormal (QR of a Gaussian matrix);
  * KS energies: occupied in [-1.0, -0.30] Ha, virtual in [0.0, 3.0] Ha;
  * AO three-centre slices T^P: symmetric, banded by exp(-|mu-nu|/32), scaled so
    that Sigma_x ~ -0.5 Ha;
  * aux Coulomb metric: identity or A A^T / N_aux + 1.
"""
from __future__ import annotations

import dataclasses

import numpy as np


@dataclasses.dataclass
class Sizes:
    """Level bookkeeping with the reference's ``ranges=default`` convention
    (upstream gwbse/gwbse.cc): rpa = all levels, qp = bse window = 0..2*homo+1."""
    n_basis: int
    n_aux: int
    homo: int
    rpamin: int = 0
    rpamax: int = -1
    qpmin: int = 0
    qpmax: int = -1
    vmin: int = 0
    cmax: int = -1

    def __post_init__(self):
        if self.rpamax < 0:
            self.rpamax = self.n_basis - 1
        if self.qpmax < 0:
            self.qpmax = min(2 * self.homo + 1, self.rpamax)
        if self.cmax < 0:
            self.cmax = min(2 * self.homo + 1, self.rpamax)

    @property
    def mmax(self): return max(self.qpmax, self.cmax)
    @property
    def mtotal(self): return self.mmax - self.rpamin + 1
    @property
    def ntotal(self): return self.rpamax - self.rpamin + 1
    @property
    def qptotal(self): return self.qpmax - self.qpmin + 1
    @property
    def n_occ(self): return self.homo - self.rpamin + 1
    @property
    def n_unocc(self): return self.rpamax - self.homo
    @property
    def vtotal(self): return self.homo - self.vmin + 1
    @property
    def ctotal(self): return self.cmax - self.homo
    @property
    def bse_size(self): return self.vtotal * self.ctotal


# named workloads: the BASELINE.json configs, by shape.  N_aux figures for the
# def2 aux sets are approximate (SURVEY.md section 8 header).
WORKLOADS = {
    "tiny": Sizes(n_basis=24, n_aux=60, homo=5),
    "ch4-svp-shape": Sizes(n_basis=34, n_aux=140, homo=4),
    "benzene-tzvp-shape": Sizes(n_basis=222, n_aux=1110, homo=20),
    "pentacene-tzvp-shape": Sizes(n_basis=766, n_aux=3830, homo=72),
    # BASELINE.json configs[2]: pentacene evGW with the contour-deformation self-energy.  CDA needs one epsilon^-1 per
    # enclosed pole and evaluation, so (as in production use of Sigma_CDA) the QP window is HOMO-4 .. LUMO+4; the RPA
    # range is all levels and the BSE window the default one (it sticks out of the QP window on both sides)
    "pentacene-tzvp-cda": Sizes(n_basis=766, n_aux=3830, homo=72, qpmin=68, qpmax=77, cmax=145),
    "c60-tzvp-shape": Sizes(n_basis=1860, n_aux=5500, homo=179),
    "synth-500": Sizes(n_basis=500, n_aux=1500, homo=49),
    "synth-1000": Sizes(n_basis=1000, n_aux=3000, homo=99),
    "synth-2000": Sizes(n_basis=2000, n_aux=6000, homo=199),
}


def make_energies(sz: Sizes, rng) -> np.ndarray:
    nocc_all = sz.homo + 1
    occ = np.sort(rng.uniform(-1.0, -0.30, nocc_all))
    virt = np.sort(rng.uniform(0.0, 3.0, sz.n_basis - nocc_all))
    return np.concatenate([occ, virt])


def make_mos(n_basis: int, rng) -> np.ndarray:
    q, _ = np.linalg.qr(rng.standard_normal((n_basis, n_basis)))
    return np.asfortranarray(q)


def target_variance(sz: Sizes) -> float:
    return 0.5 / (sz.n_occ * sz.n_aux)


def make_ao3c(sz: Sizes, rng, n_slices=None) -> np.ndarray:
    """T[P, mu, nu] symmetric in (mu, nu)."""
    nb = sz.n_basis
    n_slices = sz.n_aux if n_slices is None else n_slices
    d = np.abs(np.arange(nb)[:, None] - np.arange(nb)[None, :])
    mask = np.exp(-d / 32.0)
    t = np.sqrt(target_variance(sz) / np.mean(mask * mask))
    G = rng.standard_normal((n_slices, nb, nb))
    T = (G + np.transpose(G, (0, 2, 1))) / np.sqrt(2.0)
    return T * (mask * t)[None]


def make_aux_metric(sz: Sizes, rng, identity=False) -> np.ndarray:
    if identity:
        return np.eye(sz.n_aux)
    A = rng.standard_normal((sz.n_aux, sz.n_aux))
    return A @ A.T / sz.n_aux + np.eye(sz.n_aux)


def make_vxc(sz: Sizes, rng) -> np.ndarray:
    q = sz.qptotal
    off = 0.01 * rng.standard_normal((q, q))
    v = 0.5 * (off + off.T)
    v[np.diag_indices(q)] = -0.45 + 0.03 * rng.standard_normal(q)
    return v


def make_M_direct(sz: Sizes, rng) -> np.ndarray:
    """M[m, P, n] drawn directly (skips K1/K2), with M[m](n,P) = M[n](m,P) where
    both indices fall inside the m-window, as the real tensor has."""
    M = rng.standard_normal((sz.mtotal, sz.n_aux, sz.ntotal)) * np.sqrt(target_variance(sz))
    mt = sz.mtotal
    sq = M[:, :, :mt]
    sym = 0.5 * (sq + np.transpose(sq, (2, 1, 0))) * np.sqrt(2.0)
    M[:, :, :mt] = sym
    return np.ascontiguousarray(M)


def make_problem(name_or_sizes, seed=None, identity_metric=False):
    sz = WORKLOADS[name_or_sizes] if isinstance(name_or_sizes, str) else name_or_sizes
    rng = np.random.default_rng(20260101 + sz.n_basis if seed is None else seed)
    return {
        "sizes": sz,
        "C": make_mos(sz.n_basis, rng),
        "energies": make_energies(sz, rng),
        "ao3c": make_ao3c(sz, rng),
        "aux_coulomb": make_aux_metric(sz, rng, identity_metric),
        "vxc": make_vxc(sz, rng),
    }


def draw_window_on_device(sz: Sizes, seed, chunk=100):
    """M[m, P, n] for a tensor whose second index is restricted to the m-window (mtotal == ntotal: what the BSE operator
    needs), drawn on the GPU with torch block by block, symmetric in (m, n) as the real tensor is; same distribution as
    make_M_direct.  (v+c)^2 N_aux doubles: 61 GB at N_b 4000, where the full m x N_aux x N_b tensor would be 307 GB --
    SURVEY.md section 7 hard part 4.  Returns a torch tensor (hand its data_ptr() to TCMatrix_gwbse.set_raw_dev)."""
    import torch
    mt, na = sz.mtotal, sz.n_aux
    assert sz.ntotal == mt, "window tensor: second index range must equal the first"
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    M = torch.empty((mt, na, mt), dtype=torch.float64, device="cuda")
    scale = float(np.sqrt(target_variance(sz)))
    for i0 in range(0, mt, chunk):
        i1 = min(mt, i0 + chunk)
        for j0 in range(i0, mt, chunk):
            j1 = min(mt, j0 + chunk)
            R = torch.randn((i1 - i0, na, j1 - j0), dtype=torch.float64, device="cuda", generator=g) * scale
            if i0 == j0:
                R = (R + R.permute(2, 1, 0)) / np.sqrt(2.0)
            M[i0:i1, :, j0:j1] = R
            if i0 != j0:
                M[j0:j1, :, i0:i1] = R.permute(2, 1, 0)
    torch.cuda.synchronize()
    return M
