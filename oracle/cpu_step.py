"""A COMPLETE CPU GW-BSE step, timed stage by stage  --  TEST / BASELINE INFRASTRUCTURE ONLY (PARITY UNPINNED).

``oracle/cpu_reference.py: sampled_step`` estimates the CPU time of a C60-sized step from per-stage samples; this
module RUNS the whole step (nothing sampled, nothing scaled) on the same synthetic inputs the GPU arm uses, so that
(a) the estimate can be validated at sizes where a full step takes seconds to minutes (benzene, pentacene shape) and
(b) the "244x" of round 1 can be split into algorithm and hardware.  Two variants of the same step:

  algorithm="reference"   the reference's loop structure as recalled in SURVEY.md section 3 (upstream file per stage):
      Fill3cMO per aux function (threecenter_gwbse.cc), MultiplyRightWithAuxMatrix per slab, epsilon per occupied
      level (rpa.cc), PPM (ppm.cc), GW::SolveQP_Grid one Sigma_c evaluation at a time with an OpenMP loop over the
      levels (gw.cc / sigma_ppm.cc; C kernel cpu_gw.c: solve_qp_grid_ppm), CalcCorrelationOffDiag pair by pair
      (sigma_base.cc), BSE::SetupDirectInteractionOperator with a full tensor rotation (bse.cc), and a
      BSE_OPERATOR::matmul that REBUILDS every row block of H on every call (bse_operator.cc) under the oracle's
      Davidson solver.
  algorithm="factorised"  the algorithm of the CUDA path on the CPU: grid scan in one pass over each slab with the
      frequencies blocked in registers, off-diagonal Sigma_c as weighted-slab GEMM, eps(0) of the BSE read from the PPM
      eigenbasis, factorised BSE matmul (exchange via Mvc (Mvc^T X), direct term via the vv / cc windows).

Both produce QP and BSE energies; tests/test_cpu_step.py checks them against gwbse_oracle.run_gwbse.  GEMMs go through
numpy's BLAS on all host threads, the Sigma_c loops through OpenMP C (cpu_kernels.c, cpu_gw.c).
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

from . import cpu_reference as cr
from . import gwbse_oracle as orc

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _p(a):
    return a.ctypes.data_as(_dp)


def _lib():
    lib = cr.clib()
    if not getattr(lib, "_gw_bound", False):
        ll, i, d = C.c_longlong, C.c_int, C.c_double
        lib.solve_qp_grid_ppm.argtypes = [_dp, ll, ll, i, i, i, _dp, _dp, _dp, i, _dp, _dp, i, d, d, _dp, _ip,
                                          C.POINTER(ll)]
        lib.solve_qp_grid_ppm.restype = None
        lib.sigma_ppm_grid_batched.argtypes = [_dp, ll, ll, i, i, i, _dp, _dp, _dp, i, _dp, d, i, _dp]
        lib.sigma_ppm_grid_batched.restype = None
        lib.sigma_ppm_offdiag_all.argtypes = [_dp, ll, ll, i, i, i, _dp, _dp, _dp, i, _dp, _dp]
        lib.sigma_ppm_offdiag_all.restype = None
        lib.sigma_ppm_weighted_slab.argtypes = [_dp, ll, ll, i, i, i, _dp, _dp, _dp, i, _dp, _dp]
        lib.sigma_ppm_weighted_slab.restype = None
        lib._gw_bound = True
    return lib


class _Timer:
    def __init__(self):
        self.stage = {}

    def __call__(self, name):
        timer = self

        class _Ctx:
            def __enter__(self):
                self.t0 = time.perf_counter()

            def __exit__(self, *exc):
                timer.stage[name] = timer.stage.get(name, 0.0) + time.perf_counter() - self.t0
        return _Ctx()


class RowRebuildOperator:
    """SingletOperator_TDA as upstream BSE_OPERATOR::matmul evaluates it: for every v1 the row block
    H[(v1, c1), (v2, c2)] is rebuilt (direct term from the vv / cc windows with eps^-1, exchange from the vc window),
    multiplied with X and discarded -- 2 v^2 c^2 N_aux flops per call whatever the number of trial vectors."""

    def __init__(self, M, eps_inv, hqp, vt, ct, v0, c0, cx=2.0):
        self.vt, self.ct, self.cx = vt, ct, cx
        na = M.shape[1]
        self.Mvc = np.ascontiguousarray(np.transpose(M[v0:v0 + vt][:, :, c0:c0 + ct], (0, 2, 1)))       # [v, c, P]
        self.Mvc_flat = np.ascontiguousarray(self.Mvc.reshape(vt * ct, na).T)                           # [P, (v c)]
        self.Mvv_s = np.ascontiguousarray(np.transpose(M[v0:v0 + vt][:, :, v0:v0 + vt], (0, 2, 1)) * eps_inv)   # [v1, v2, P]
        self.Mcc_flat = np.ascontiguousarray(np.transpose(M[c0:c0 + ct][:, :, c0:c0 + ct], (1, 0, 2)).reshape(na, ct * ct))
        self.hqp = hqp
        self.calls = 0

    def rows(self): return self.vt * self.ct

    def diagonal(self):
        vt, ct = self.vt, self.ct
        hq = np.diag(self.hqp)
        d = hq[vt:][None, :] - hq[:vt][:, None]
        d = d + self.cx * np.einsum("vcp,vcp->vc", self.Mvc, self.Mvc)
        dv = np.einsum("vvp->vp", self.Mvv_s)
        dc = self.Mcc_flat.reshape(-1, ct, ct)[:, np.arange(ct), np.arange(ct)].T                       # [c, P]
        return (d - dv @ dc.T).reshape(-1)

    def matmul(self, X):
        vt, ct = self.vt, self.ct
        self.calls += 1
        k = X.shape[1]
        X4 = X.reshape(vt, ct, k)
        Y = np.einsum("cd,vdk->vck", self.hqp[vt:, vt:], X4) - np.einsum("vw,wck->vck", self.hqp[:vt, :vt], X4)
        for v1 in range(vt):
            Hd = (self.Mvv_s[v1] @ self.Mcc_flat).reshape(vt, ct, ct)            # [v2, c1, c2]
            Hx = self.Mvc[v1] @ self.Mvc_flat                                    # [c1, (v2 c2)]
            H = self.cx * Hx - np.transpose(Hd, (1, 0, 2)).reshape(ct, vt * ct)
            Y[v1] += H @ X
        return Y.reshape(vt * ct, k)


class FactorisedOperator(RowRebuildOperator):
    """The same operator applied through the RI factorisation (no H row is ever formed)."""

    def __init__(self, M, eps_inv, hqp, vt, ct, v0, c0, cx=2.0):
        super().__init__(M, eps_inv, hqp, vt, ct, v0, c0, cx)
        na = M.shape[1]
        self.Mcc_pc = np.ascontiguousarray(np.transpose(M[c0:c0 + ct][:, :, c0:c0 + ct], (1, 0, 2)))    # [P, c1, c2]
        self.Mvv_pv = np.ascontiguousarray(np.transpose(self.Mvv_s, (2, 0, 1)))                        # [P, v1, v2]
        self.Mvc_mat = self.Mvc.reshape(vt * ct, na)

    def matmul(self, X):
        vt, ct = self.vt, self.ct
        self.calls += 1
        k = X.shape[1]
        X4 = X.reshape(vt, ct, k)
        Y = np.einsum("cd,vdk->vck", self.hqp[vt:, vt:], X4) - np.einsum("vw,wck->vck", self.hqp[:vt, :vt], X4)
        Y = Y.reshape(vt * ct, k) + self.cx * (self.Mvc_mat @ (self.Mvc_mat.T @ X))
        Xt = np.ascontiguousarray(np.transpose(X4, (1, 0, 2)).reshape(ct, vt * k))                      # [c2, (v2 k)]
        U = (self.Mcc_pc.reshape(-1, ct) @ Xt).reshape(-1, ct, vt, k)                                   # [P, c1, v2, k]
        Yd = np.einsum("pvw,pcwk->vck", self.Mvv_pv, U, optimize=True)
        return Y - Yd.reshape(vt * ct, k)


def run_step(prob, algorithm="reference", nmax=10, grid_steps=1001, grid_spacing=0.01, davidson_tolerance="normal"):
    """One complete G0W0 (PPM) + BSE (TDA singlets) step on the CPU.  Returns wall seconds, per-stage seconds, energies."""
    assert algorithm in ("reference", "factorised")
    threads = cr.use_all_host_threads()
    lib = _lib()
    T = _Timer()
    sz = prob["sizes"]
    C_mo, e_dft, vxc, V, ao = prob["C"], prob["energies"], prob["vxc"], prob["aux_coulomb"], prob["ao3c"]
    na, mt, nt, nocc, q = sz.n_aux, sz.mtotal, sz.ntotal, sz.n_occ, sz.qptotal
    q0 = sz.qpmin - sz.rpamin
    t_start = time.perf_counter()

    with T("fill"):                                      # TCMatrix_gwbse::Fill3cMO
        Cm = np.ascontiguousarray(C_mo[:, sz.rpamin:sz.rpamin + mt])
        Cn = np.ascontiguousarray(C_mo[:, sz.rpamin:sz.rpamax + 1])
        M = np.empty((mt, na, nt))
        for P in range(na):
            M[:, P, :] = (Cn.T @ (ao[P] @ Cm)).T
    with T("metric"):                                    # Pseudo_InvSqrt_GWBSE + MultiplyRightWithAuxMatrix
        lam, U = np.linalg.eigh(V)
        R = (U / np.sqrt(lam)) @ U.T
        for m in range(mt):
            M[m] = R.T @ M[m]

    e_rpa = np.array(e_dft[sz.rpamin:sz.rpamax + 1])

    def epsilon(w, imag):                                # RPA::calculate_epsilon, loop over occupied levels
        dE = e_rpa[nocc:][None, :] - e_rpa[:nocc][:, None]
        if imag:
            d = 4.0 * dE / (dE * dE + w * w)
        else:
            eta2 = 1e-6
            d = 2.0 * ((dE - w) / ((dE - w) ** 2 + eta2) + (dE + w) / ((dE + w) ** 2 + eta2))
        eps = np.eye(na)
        for m in range(nocc):
            A = M[m][:, nocc:]
            eps += (A * d[m]) @ A.T
        return eps

    with T("epsilon"):
        eps0 = epsilon(0.0, False)
        eps1 = epsilon(0.5, True)
    with T("ppm"):                                       # PPM::PPM_construct_parameters + rotation into its eigenbasis
        lam0, phi = np.linalg.eigh(eps0)
        weight = 1.0 - 1.0 / lam0
        x = np.diag(np.linalg.inv(phi.T @ eps1 @ phi)) - 1.0
        freq = np.where(weight < 1e-5, 0.5, np.sqrt(np.abs(-x / (x + np.where(weight < 1e-5, 1.0, weight)) * 0.25)))
        weight = np.where(weight < 1e-5, 0.0, weight)
        fac = np.where(weight < 1e-9, 0.0, 0.5 * weight * freq)
        for m in range(mt):
            M[m] = phi.T @ M[m]
    with T("sigma_x"):                                   # Sigma_base::CalcExchangeMatrix
        B = M[q0:q0 + q][:, :, :nocc].reshape(q, -1)
        sigma_x = -(B @ B.T)

    slabs = M[q0:]                                       # gw level l at slabs[l]
    intercept = e_dft[sz.qpmin:sz.qpmin + q] + np.diag(sigma_x) - np.diag(vxc)
    f0 = np.array(e_dft[sz.qpmin:sz.qpmin + q])
    args = (_p(slabs), na * nt, nt, nt, na, nocc, _p(e_rpa), _p(freq), _p(fac))
    with T("sigma_c"):                                   # GW::SolveQP (grid) + final diagonal
        qp = np.empty(q)
        conv = np.zeros(q, dtype=np.int32)
        nev = C.c_longlong(0)
        if algorithm == "reference":
            lib.solve_qp_grid_ppm(*args, q, _p(intercept), _p(f0), grid_steps, grid_spacing, 1e-5, _p(qp),
                                  conv.ctypes.data_as(_ip), C.byref(nev))
        else:
            # batched scan of all grid points, then the bracketed roots are refined with the scalar routine on a
            # 3-point grid around each sign change (same bisection, same root choice)
            rng_ = grid_spacing * (grid_steps - 1) / 2.0
            om0 = f0 - rng_
            vals = np.empty((q, grid_steps))
            lib.sigma_ppm_grid_batched(*args, q, _p(om0), grid_spacing, grid_steps, _p(vals))
            grid_w = om0[:, None] + grid_spacing * np.arange(grid_steps)[None, :]
            tvals = vals + intercept[:, None] - grid_w
            best = np.full(q, np.inf)
            for l in range(q):
                sign = np.nonzero(tvals[l, :-1] * tvals[l, 1:] < 0.0)[0]
                found = False
                for j in sign:
                    lo, hi, flo = grid_w[l, j], grid_w[l, j + 1], tvals[l, j]
                    while True:
                        c = 0.5 * (lo + hi)
                        if abs(hi - lo) < 1e-5:
                            break
                        yc = cr.sigma_ppm_diag(slabs[l], nocc, e_rpa, freq, fac, [c])[0] + intercept[l] - c
                        if abs(yc) < 1e-5:
                            break
                        if yc * flo > 0:
                            lo, flo = c, yc
                        else:
                            hi = c
                    _, der = cr.sigma_ppm_diag(slabs[l], nocc, e_rpa, freq, fac, [c], deriv=True)
                    if abs(der[0] - 1.0) < best[l]:
                        best[l], qp[l], found = abs(der[0] - 1.0), c, True
                conv[l] = int(found)
                if not found:
                    s, der = cr.sigma_ppm_diag(slabs[l], nocc, e_rpa, freq, fac, [f0[l]], deriv=True)
                    Z = 1.0 - der[0]
                    qp[l] = f0[l] + (intercept[l] - f0[l] + s[0]) / Z if abs(Z) > 1e-9 else f0[l]
        sigma_c_diag = np.array([cr.sigma_ppm_diag(slabs[l], nocc, e_rpa, freq, fac, [qp[l]])[0] for l in range(q)])
    qp_pert = e_dft[sz.qpmin:sz.qpmin + q] + np.diag(sigma_x) + sigma_c_diag - np.diag(vxc)
    with T("offdiag"):                                   # Sigma_base::CalcCorrelationOffDiag at the QP energies
        if algorithm == "reference":
            off = np.empty((q, q))
            lib.sigma_ppm_offdiag_all(*args, q, _p(np.ascontiguousarray(qp_pert)), _p(off))
        else:
            W = np.empty((q, na, nt))
            lib.sigma_ppm_weighted_slab(*args, q, _p(np.ascontiguousarray(qp_pert)), _p(W))
            S = W.reshape(q, -1) @ slabs[:q].reshape(q, -1).T
            off = 0.5 * (S + S.T)
            off[np.diag_indices(q)] = 0.0
    hqp = sigma_x + off - vxc
    hqp[np.diag_indices(q)] = qp_pert

    with T("bse_setup"):                                 # BSE::SetupDirectInteractionOperator
        if algorithm == "reference":
            lamb, Ub = np.linalg.eigh(epsilon(0.0, False))
            for m in range(mt):
                M[m] = Ub.T @ M[m]
        else:
            lamb = lam0                                  # the tensor already is in the eigenbasis of eps(0)
        eps_inv = np.where(lamb > 1e-8, 1.0 / np.where(lamb > 1e-8, lamb, 1.0), 0.0)
        vt, ct = sz.vtotal, sz.ctotal
        assert sz.vmin == sz.qpmin and sz.cmax == sz.qpmax, "default ranges (BSE window == QP window) only"
        cls = RowRebuildOperator if algorithm == "reference" else FactorisedOperator
        op = cls(M, eps_inv, hqp, vt, ct, sz.vmin - sz.rpamin, sz.homo + 1 - sz.rpamin)
    with T("davidson"):
        ds = orc.DavidsonSolver()
        ds.set_tolerance(davidson_tolerance)
        ds.set_max_search_space(10 * nmax)
        ds.solve(op, nmax)
    total = time.perf_counter() - t_start
    return {"seconds": total, "stage_seconds": {k: round(v, 4) for k, v in T.stage.items()}, "threads": threads,
            "algorithm": algorithm, "qp": qp_pert, "qp_converged": conv, "singlets": ds.eigenvalues(),
            "davidson_iterations": ds.num_iterations(), "matmul_calls": op.calls,
            "sigma_c_evaluations": int(nev.value)}
