"""Real-molecule inputs for the GW-BSE path without a quantum-chemistry package.

The hot path takes MO coefficients, orbital energies, AO three-centre integrals (mu nu|P), the aux Coulomb metric
(P|Q) and Vxc.  Upstream they come from XTP's integral layer (libint2: `aomatrices/`, `threecenter.cc`) and a DFT run
(`dftengine.cc` or ORCA) -- none of which exists in this image.  This module is a small, plain-numpy stand-in so that
the path can be exercised on physically meaningful numbers: Cartesian Gaussian integrals by the McMurchie-Davidson
scheme (Hermite expansion coefficients E_t^{ij}, Hermite Coulomb integrals R_tuv, Boys function through 1F1), the
STO-3G basis for H, C, N, O, an even-tempered auxiliary basis, and a restricted Hartree-Fock SCF.  G0W0@HF needs no
exchange-correlation functional: with ScaHFX = 1 the reference's QP equation reduces to e_QP = e_HF + Sigma_c(e_QP)
and Vxc = 0 (upstream gw.cc: Sigma_x is scaled by 1 - ScaHFX).

Host-side INPUT generation only (tests, examples): nothing here is on the product's compute path, and it is sized for
molecules of a few atoms (pure-Python loops over primitive quartets).
"""
from __future__ import annotations

import itertools
import math

import numpy as np
from scipy.special import hyp1f1

# STO-3G (Hehre, Stewart, Pople 1969): exponents and contraction coefficients of normalised primitives
_STO3G = {
    "H": [("s", [3.42525091, 0.62391373, 0.16885540], [0.15432897, 0.53532814, 0.44463454])],
    "C": [("s", [71.6168370, 13.0450960, 3.5305122], [0.15432897, 0.53532814, 0.44463454]),
          ("s", [2.9412494, 0.6834831, 0.2222899], [-0.09996723, 0.39951283, 0.70011547]),
          ("p", [2.9412494, 0.6834831, 0.2222899], [0.15591627, 0.60768372, 0.39195739])],
    "N": [("s", [99.1061690, 18.0523120, 4.8856602], [0.15432897, 0.53532814, 0.44463454]),
          ("s", [3.7804559, 0.8784966, 0.2857144], [-0.09996723, 0.39951283, 0.70011547]),
          ("p", [3.7804559, 0.8784966, 0.2857144], [0.15591627, 0.60768372, 0.39195739])],
    "O": [("s", [130.7093200, 23.8088610, 6.4436083], [0.15432897, 0.53532814, 0.44463454]),
          ("s", [5.0331513, 1.1695961, 0.3803890], [-0.09996723, 0.39951283, 0.70011547]),
          ("p", [5.0331513, 1.1695961, 0.3803890], [0.15591627, 0.60768372, 0.39195739])],
}
_Z = {"H": 1, "C": 6, "N": 7, "O": 8}
_L = {"s": 0, "p": 1, "d": 2}


def _cart(l):
    """Cartesian powers (lx, ly, lz) of shell l in the usual order (x, y, z / xx, xy, xz, yy, yz, zz)."""
    return [(lx, ly, l - lx - ly) for lx in range(l, -1, -1) for ly in range(l - lx, -1, -1)]


def _dfact(n):
    return 1.0 if n <= 0 else n * _dfact(n - 2)


def _norm(a, lmn):
    l, m, n = lmn
    return ((2 * a / math.pi) ** 0.75 * (4 * a) ** ((l + m + n) / 2.0)
            / math.sqrt(_dfact(2 * l - 1) * _dfact(2 * m - 1) * _dfact(2 * n - 1)))


class BasisFunction:
    """Contracted Cartesian Gaussian: sum_k c_k N_k x^l y^m z^n exp(-a_k r^2) centred at `center`."""

    def __init__(self, center, lmn, exps, coefs):
        self.center = np.asarray(center, dtype=float)
        self.lmn = tuple(lmn)
        self.exps = [float(a) for a in exps]
        self.coefs = [float(c) * _norm(a, lmn) for a, c in zip(exps, coefs)]
        # normalise the contraction (matters for uncontracted aux functions with lx=2 etc. too)
        s = sum(ci * cj * _overlap_prim(ai, self.lmn, self.center, aj, self.lmn, self.center)
                for ai, ci in zip(self.exps, self.coefs) for aj, cj in zip(self.exps, self.coefs))
        self.coefs = [c / math.sqrt(s) for c in self.coefs]


def _E(i, j, t, Qx, a, b):
    """Hermite expansion coefficient E_t^{ij} of x_A^i x_B^j exp(-a x_A^2 - b x_B^2) (McMurchie-Davidson)."""
    p = a + b
    q = a * b / p
    if t < 0 or t > i + j:
        return 0.0
    if i == j == t == 0:
        return math.exp(-q * Qx * Qx)
    if j == 0:
        return (_E(i - 1, j, t - 1, Qx, a, b) / (2 * p) - q * Qx / a * _E(i - 1, j, t, Qx, a, b)
                + (t + 1) * _E(i - 1, j, t + 1, Qx, a, b))
    return (_E(i, j - 1, t - 1, Qx, a, b) / (2 * p) + q * Qx / b * _E(i, j - 1, t, Qx, a, b)
            + (t + 1) * _E(i, j - 1, t + 1, Qx, a, b))


def _overlap_prim(a, lmn1, A, b, lmn2, B):
    p = a + b
    return (math.pi / p) ** 1.5 * math.prod(_E(lmn1[k], lmn2[k], 0, A[k] - B[k], a, b) for k in range(3))


def _kinetic_prim(a, lmn1, A, b, lmn2, B):
    l2, m2, n2 = lmn2
    t0 = b * (2 * (l2 + m2 + n2) + 3) * _overlap_prim(a, lmn1, A, b, lmn2, B)
    t1 = -2 * b * b * (_overlap_prim(a, lmn1, A, b, (l2 + 2, m2, n2), B) + _overlap_prim(a, lmn1, A, b, (l2, m2 + 2, n2), B)
                       + _overlap_prim(a, lmn1, A, b, (l2, m2, n2 + 2), B))
    t2 = -0.5 * (l2 * (l2 - 1) * _overlap_prim(a, lmn1, A, b, (l2 - 2, m2, n2), B)
                 + m2 * (m2 - 1) * _overlap_prim(a, lmn1, A, b, (l2, m2 - 2, n2), B)
                 + n2 * (n2 - 1) * _overlap_prim(a, lmn1, A, b, (l2, m2, n2 - 2), B))
    return t0 + t1 + t2


def _boys(n, x):
    return hyp1f1(n + 0.5, n + 1.5, -x) / (2.0 * n + 1.0)


def _R(t, u, v, n, p, PC, cache):
    """Hermite Coulomb integral R^n_{tuv}(p, PC)."""
    key = (t, u, v, n)
    if key in cache:
        return cache[key]
    x, y, z = PC
    if t == u == v == 0:
        val = (-2.0 * p) ** n * _boys(n, p * (x * x + y * y + z * z))
    elif t == u == 0:
        val = (v - 1) * _R(t, u, v - 2, n + 1, p, PC, cache) if v > 1 else 0.0
        val += z * _R(t, u, v - 1, n + 1, p, PC, cache)
    elif t == 0:
        val = (u - 1) * _R(t, u - 2, v, n + 1, p, PC, cache) if u > 1 else 0.0
        val += y * _R(t, u - 1, v, n + 1, p, PC, cache)
    else:
        val = (t - 1) * _R(t - 2, u, v, n + 1, p, PC, cache) if t > 1 else 0.0
        val += x * _R(t - 1, u, v, n + 1, p, PC, cache)
    cache[key] = val
    return val


def _hermite_pair(a, lmn1, A, b, lmn2, B):
    """[(t, u, v, E_t E_u E_v)] of a primitive product, its total exponent and centre."""
    p = a + b
    P = (a * A + b * B) / p
    Ex = [_E(lmn1[0], lmn2[0], t, A[0] - B[0], a, b) for t in range(lmn1[0] + lmn2[0] + 1)]
    Ey = [_E(lmn1[1], lmn2[1], t, A[1] - B[1], a, b) for t in range(lmn1[1] + lmn2[1] + 1)]
    Ez = [_E(lmn1[2], lmn2[2], t, A[2] - B[2], a, b) for t in range(lmn1[2] + lmn2[2] + 1)]
    terms = [(t, u, v, Ex[t] * Ey[u] * Ez[v]) for t in range(len(Ex)) for u in range(len(Ey)) for v in range(len(Ez))
             if Ex[t] * Ey[u] * Ez[v] != 0.0]
    return terms, p, P


def _hermite_single(a, lmn, A):
    """the same for one primitive (a product with an s function of exponent 0)"""
    return _hermite_pair(a, lmn, A, 0.0, (0, 0, 0), A)


def _coulomb_hermite(bra, ket):
    """Coulomb interaction of two Hermite charge distributions."""
    tb, p, P = bra
    tk, q, Q = ket
    alpha = p * q / (p + q)
    cache = {}
    PQ = P - Q
    val = 0.0
    for t, u, v, eb in tb:
        for tt, uu, vv, ek in tk:
            val += eb * ek * (-1) ** (tt + uu + vv) * _R(t + tt, u + uu, v + vv, 0, alpha, PQ, cache)
    return val * 2.0 * math.pi ** 2.5 / (p * q * math.sqrt(p + q))


class _Distributions:
    """Hermite expansions of every primitive pair of a basis (or every primitive of an aux basis), built once."""

    def __init__(self, basis, pairs=True):
        self.items = {}
        n = len(basis)
        if pairs:
            for i in range(n):
                for j in range(i + 1):
                    self.items[(i, j)] = [(ci * cj, _hermite_pair(ai, basis[i].lmn, basis[i].center, aj, basis[j].lmn,
                                                                  basis[j].center))
                                          for ai, ci in zip(basis[i].exps, basis[i].coefs)
                                          for aj, cj in zip(basis[j].exps, basis[j].coefs)]
        else:
            for i in range(n):
                self.items[i] = [(ci, _hermite_single(ai, basis[i].lmn, basis[i].center))
                                 for ai, ci in zip(basis[i].exps, basis[i].coefs)]


def _contract(d1, d2):
    return sum(c1 * c2 * _coulomb_hermite(h1, h2) for c1, h1 in d1 for c2, h2 in d2)


class Molecule:
    def __init__(self, atoms):
        """atoms: [(symbol, (x, y, z) in bohr)]"""
        self.atoms = [(s, np.asarray(r, dtype=float)) for s, r in atoms]

    @property
    def n_electrons(self):
        return sum(_Z[s] for s, _ in self.atoms)

    def nuclear_repulsion(self):
        return sum(_Z[a] * _Z[b] / np.linalg.norm(ra - rb) for (a, ra), (b, rb) in itertools.combinations(self.atoms, 2))

    def sto3g(self):
        basis = []
        for sym, r in self.atoms:
            for kind, exps, coefs in _STO3G[sym]:
                for lmn in _cart(_L[kind]):
                    basis.append(BasisFunction(r, lmn, exps, coefs))
        return basis

    def even_tempered_aux(self, lmax=2, n_per_l=(10, 6, 3), ratio=2.6, amin=(0.25, 0.35, 0.6)):
        """Uncontracted even-tempered auxiliary functions on every atom (Cartesian s, p, d).  Heavy atoms get the full
        set, hydrogen one function less per angular momentum.  Stand-in for the def2 aux sets of the reference; with the
        defaults the RI error of the STO-3G four-index integrals of water is 9e-5 Ha (tests/test_molecule_inputs.py)."""
        aux = []
        for sym, r in self.atoms:
            for l in range(lmax + 1):
                cnt = n_per_l[l] - (1 if sym == "H" else 0)
                for k in range(max(cnt, 0)):
                    a = amin[l] * ratio ** k
                    for lmn in _cart(l):
                        aux.append(BasisFunction(r, lmn, [a], [1.0]))
        return aux


def one_electron(mol, basis):
    n = len(basis)
    S, T, V = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1):
            bi, bj = basis[i], basis[j]
            s = t = v = 0.0
            for ai, ci in zip(bi.exps, bi.coefs):
                for aj, cj in zip(bj.exps, bj.coefs):
                    s += ci * cj * _overlap_prim(ai, bi.lmn, bi.center, aj, bj.lmn, bj.center)
                    t += ci * cj * _kinetic_prim(ai, bi.lmn, bi.center, aj, bj.lmn, bj.center)
                    terms, p, P = _hermite_pair(ai, bi.lmn, bi.center, aj, bj.lmn, bj.center)
                    for sym, C in mol.atoms:
                        cache = {}
                        v -= ci * cj * _Z[sym] * 2.0 * math.pi / p * sum(e * _R(tt, u, w, 0, p, P - C, cache)
                                                                         for tt, u, w, e in terms)
            S[i, j] = S[j, i] = s
            T[i, j] = T[j, i] = t
            V[i, j] = V[j, i] = v
    return S, T, V


def dipole_matrices(basis):
    """<mu| r_k |nu> about the origin, k = x, y, z (upstream AODipole): x_B^(j+1) + B_x x_B^j on the ket side."""
    n = len(basis)
    D = np.zeros((3, n, n))
    for i in range(n):
        for j in range(n):
            bi, bj = basis[i], basis[j]
            for k in range(3):
                up = list(bj.lmn)
                up[k] += 1
                val = 0.0
                for ai, ci in zip(bi.exps, bi.coefs):
                    for aj, cj in zip(bj.exps, bj.coefs):
                        val += ci * cj * (_overlap_prim(ai, bi.lmn, bi.center, aj, tuple(up), bj.center)
                                          + bj.center[k] * _overlap_prim(ai, bi.lmn, bi.center, aj, bj.lmn, bj.center))
                D[k, i, j] = val
    return 0.5 * (D + np.transpose(D, (0, 2, 1)))


def eri_four_center(basis):
    """(mu nu|la si), full array with the eightfold symmetry filled in (small molecules only)."""
    n = len(basis)
    dist = _Distributions(basis)
    pairs = [(i, j) for i in range(n) for j in range(i + 1)]
    G = np.zeros((n, n, n, n))
    for a, (i, j) in enumerate(pairs):
        for (k, l) in pairs[:a + 1]:
            v = _contract(dist.items[(i, j)], dist.items[(k, l)])
            for x, y in ((i, j), (j, i)):
                for z, w in ((k, l), (l, k)):
                    G[x, y, z, w] = G[z, w, x, y] = v
    return G


def eri_three_center(basis, aux):
    """T[P, mu, nu] = (mu nu|P): the slices TCMatrix_gwbse::Fill3cMO consumes (upstream threecenter.cc)."""
    n, na = len(basis), len(aux)
    dist, adist = _Distributions(basis), _Distributions(aux, pairs=False)
    T = np.zeros((na, n, n))
    for (i, j), d in dist.items.items():
        for P in range(na):
            T[P, i, j] = T[P, j, i] = _contract(d, adist.items[P])
    return T


def eri_two_center(aux):
    """V[P, Q] = (P|Q): the aux Coulomb metric (upstream AOCoulomb::Fill)."""
    na = len(aux)
    adist = _Distributions(aux, pairs=False)
    V = np.zeros((na, na))
    for P in range(na):
        for Q in range(P + 1):
            V[P, Q] = V[Q, P] = _contract(adist.items[P], adist.items[Q])
    return V


def rhf(mol, basis, eri=None, max_iter=200, tol=1e-10):
    """Restricted Hartree-Fock with DIIS.  Returns dict(energy, C, eps, n_occ, S, hcore, eri)."""
    S, T, V = one_electron(mol, basis)
    G = eri_four_center(basis) if eri is None else eri
    H = T + V
    nocc = mol.n_electrons // 2
    lam, U = np.linalg.eigh(S)
    X = U / np.sqrt(lam)
    eps, Cp = np.linalg.eigh(X.T @ H @ X)
    C = X @ Cp
    focks, errs = [], []
    e_old = 0.0
    for _ in range(max_iter):
        D = 2.0 * C[:, :nocc] @ C[:, :nocc].T
        F = H + np.einsum("ls,mnls->mn", D, G) - 0.5 * np.einsum("ls,mlns->mn", D, G)
        e = 0.5 * np.sum(D * (H + F)) + mol.nuclear_repulsion()
        err = X.T @ (F @ D @ S - S @ D @ F) @ X
        focks.append(F)
        errs.append(err)
        focks, errs = focks[-8:], errs[-8:]
        if len(focks) > 1:
            m = len(focks)
            B = -np.ones((m + 1, m + 1))
            B[m, m] = 0.0
            for a in range(m):
                for b in range(m):
                    B[a, b] = np.sum(errs[a] * errs[b])
            rhs = np.zeros(m + 1)
            rhs[m] = -1.0
            try:
                c = np.linalg.solve(B, rhs)[:m]
                F = sum(ci * Fi for ci, Fi in zip(c, focks))
            except np.linalg.LinAlgError:
                pass
        eps, Cp = np.linalg.eigh(X.T @ F @ X)
        C = X @ Cp
        if abs(e - e_old) < tol and np.abs(err).max() < 1e-8:
            break
        e_old = e
    # fix the sign convention of every MO (largest coefficient positive) so that fixtures are reproducible
    for k in range(C.shape[1]):
        if C[np.argmax(np.abs(C[:, k])), k] < 0:
            C[:, k] = -C[:, k]
    return {"energy": float(e), "C": np.asfortranarray(C), "eps": eps, "n_occ": nocc, "S": S, "hcore": H, "eri": G}


def water():
    """H2O at the geometry of the usual STO-3G teaching example (bohr): E_nuc = 8.002367061810, E_RHF = -74.942079928."""
    return Molecule([("O", (0.000000000000, -0.143225816552, 0.000000000000)),
                     ("H", (1.638036840407, 1.136548822547, 0.000000000000)),
                     ("H", (-1.638036840407, 1.136548822547, 0.000000000000))])


def gwbse_inputs(mol):
    """Everything the path needs for G0W0@HF + BSE on `mol` with STO-3G and the even-tempered aux basis."""
    basis, aux = mol.sto3g(), mol.even_tempered_aux()
    scf = rhf(mol, basis)
    return {"C": scf["C"], "energies": scf["eps"], "n_occ": scf["n_occ"], "ao3c": eri_three_center(basis, aux),
            "aux_coulomb": eri_two_center(aux), "ao_dipoles": dipole_matrices(basis), "scf": scf,
            "n_basis": len(basis), "n_aux": len(aux)}
