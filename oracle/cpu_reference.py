"""Timed CPU baseline of the GW-BSE step  --  TEST / BASELINE INFRASTRUCTURE ONLY (PARITY UNPINNED).

votca/xtp cannot be built here (``/root/reference/README.md:1`` is a redirect stub; Eigen, libint2, libxc, HDF5 are
absent), so the CPU number printed beside the GPU number is this restatement, labelled everywhere as
**"restated CPU baseline, not votca/xtp binaries"** (BASELINE.md section 4).  It mirrors the reference's algorithmic
structure as recalled in SURVEY.md section 3 (upstream files named per stage):

  fill      TCMatrix_gwbse::Fill3cMO        per aux function  C_n^T (T_P C_m)            threecenter_gwbse.cc
  metric    Pseudo_InvSqrt_GWBSE + MultiplyRightWithAuxMatrix  per slab  M[m] A          aocoulomb.cc / threecenter_gwbse.cc
  epsilon   RPA::calculate_epsilon          per occupied level  A_m^T diag(d) A_m        rpa.cc
  ppm       PPM::PPM_construct_parameters   eigh + inverse + 2 GEMM, then M <- M Phi     ppm.cc / sigma_ppm.cc
  sigma_x   Sigma_base::CalcExchangeMatrix                                               sigma_base.cc
  sigma_c   GW::SolveQP_Grid -> Sigma_PPM::CalcCorrelationDiagElement one (level, w) at a time   gw.cc / sigma_ppm.cc
  offdiag   Sigma_base::CalcCorrelationOffDiag                                           sigma_base.cc
  bse_setup BSE::SetupDirectInteractionOperator  epsilon(0), eigh, M <- M U              bse.cc
  davidson  BSE_OPERATOR::matmul rebuilding every row block of H on every call
            (2 v^2 c^2 N_aux flops per call, independent of the number of trial vectors)  bse_operator.cc

GEMMs run through numpy's OpenBLAS on all host threads, the Sigma_c loops through the OpenMP C kernels of
``cpu_kernels.c``.  A full C60-sized step would take ~20 minutes of CPU, so each stage is timed on a BOUNDED SAMPLE
of its independent units (aux functions, slabs, occupied levels, frequencies, H row blocks) and scaled by the unit
count; the dense eigensolver / inverse are timed at a reduced size and scaled cubically.  ``describe`` in the result
says exactly what was sampled.  Small workloads are run completely (``full_step``) through ``gwbse_oracle.run_gwbse``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def clib():
    """liboracle_cpu.so, compiled on first use if the prebuilt file is missing."""
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle_cpu.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
        lib = C.CDLL(path)
        dp = C.POINTER(C.c_double)
        lib.sigma_ppm_diag.argtypes = [dp, C.c_longlong, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, C.c_int, dp, dp]
        lib.sigma_ppm_diag.restype = None
        lib.unpack_symmetric.argtypes = [dp, C.c_int, dp, C.c_longlong]
        lib.unpack_symmetric.restype = None
        lib.oracle_num_threads.restype = C.c_int
        lib.oracle_set_num_threads.argtypes = [C.c_int]
        lib.oracle_set_num_threads.restype = None
        _LIB = lib
    return _LIB


_THREADS_SET = False


def use_all_host_threads():
    """Make OpenBLAS (numpy) and the OpenMP C kernels use every core this process may run on, whatever
    OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to its ranks).  Returns the thread count in use."""
    global _THREADS_SET
    n = host_threads()
    if not _THREADS_SET:
        clib().oracle_set_num_threads(n)
        try:
            import threadpoolctl
            threadpoolctl.threadpool_limits(limits=n)      # process-wide when not used as a context manager
        except Exception:  # noqa: BLE001  (threadpoolctl missing: numpy keeps its start-up thread count)
            pass
        _THREADS_SET = True
    return n


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def host_threads():
    return len(os.sched_getaffinity(0))


def sigma_ppm_diag(slab, n_occ, energies, ppm_freq, ppm_fac, omegas, deriv=False):
    """slab[P, m] (C-contiguous) -> Sigma_c(level, w) for every w (and d/dw)."""
    slab = np.ascontiguousarray(slab, dtype=np.float64)
    om = np.ascontiguousarray(omegas, dtype=np.float64)
    out = np.empty(len(om))
    der = np.empty(len(om)) if deriv else None
    e = np.ascontiguousarray(energies, dtype=np.float64)
    f = np.ascontiguousarray(ppm_freq, dtype=np.float64)
    g = np.ascontiguousarray(ppm_fac, dtype=np.float64)
    clib().sigma_ppm_diag(_dp(slab), slab.shape[1], slab.shape[1], slab.shape[0], int(n_occ), _dp(e), _dp(f), _dp(g),
                          _dp(om), len(om), _dp(out), _dp(der) if deriv else None)
    return (out, der) if deriv else out


def unpack_symmetric(packed, n):
    full = np.empty((n, n), order="F")
    p = np.ascontiguousarray(packed, dtype=np.float64)
    clib().unpack_symmetric(_dp(p), int(n), _dp(full), n)
    return full


def _timeit(fn, min_reps=1):
    fn()                      # warm (page faults, BLAS thread spin-up)
    t0 = time.perf_counter()
    for _ in range(min_reps):
        fn()
    return (time.perf_counter() - t0) / min_reps


def sampled_step(sz, davidson_matmul_calls=12, grid_steps=1001, triplets=False, scale=1.0, seed=7):
    """Estimated seconds per molecule for workload ``sz`` (xtp_b200.synth.Sizes), reference-structure CPU.
    ``scale`` multiplies every sample size (1.0 ~ 15-25 s of CPU work on 16 cores at C60 size)."""
    use_all_host_threads()
    rng = np.random.default_rng(seed)
    nb, naux = sz.n_basis, sz.n_aux
    m, n, o = sz.mtotal, sz.ntotal, sz.n_occ
    u = n - o
    q, vt, ct = sz.qptotal, sz.vtotal, sz.ctotal
    stage, desc = {}, {}

    def cnt(base, total):
        return int(max(1, min(total, round(base * scale))))

    # --- fill: per aux function C_n^T (T_P C_m)
    ns = cnt(6, naux)
    Cm = rng.standard_normal((nb, m))
    Cn = rng.standard_normal((nb, n))
    T = rng.standard_normal((nb, nb))
    T = T + T.T
    t = _timeit(lambda: [Cn.T @ (T @ Cm) for _ in range(ns)])
    stage["fill"] = t * naux / ns
    desc["fill"] = f"{ns} of {naux} aux functions"
    del Cm, Cn, T

    # --- dense solver pieces at reduced size, cubic scaling
    ne = min(naux, 1536)
    S = rng.standard_normal((ne, ne))
    S = S @ S.T / ne + np.eye(ne)
    cubic = (naux / ne) ** 3
    t_eigh = _timeit(lambda: np.linalg.eigh(S)) * cubic
    t_inv = _timeit(lambda: np.linalg.inv(S)) * cubic
    t_gemm = _timeit(lambda: S @ S) * cubic
    desc["solver"] = f"eigh/inverse/GEMM at n={ne}, scaled by (N_aux/n)^3"
    del S

    # --- aux rotation: per slab M[m] (n x naux) @ A (naux x naux)
    nr = cnt(2, m)
    A = rng.standard_normal((naux, naux))
    slab = rng.standard_normal((n, naux))
    t_rot_slab = _timeit(lambda: [slab @ A for _ in range(nr)]) / nr
    desc["rotation"] = f"{nr} of {m} slabs"
    stage["metric"] = t_eigh + t_gemm + t_rot_slab * m      # one eigh + U s U^T + rotation
    # --- epsilon: per occupied level A_m^T diag(d) A_m
    nE = cnt(2, o)
    Am = np.ascontiguousarray(slab[:u])
    d = rng.random(u)
    t_eps_level = _timeit(lambda: [(Am.T * d) @ Am for _ in range(nE)]) / nE
    t_eps = t_eps_level * o
    desc["epsilon"] = f"{nE} of {o} occupied levels per frequency; 3 frequencies per step (PPM 2, BSE 1)"
    stage["epsilon"] = 3 * t_eps
    stage["ppm"] = t_eigh + t_inv + 2 * t_gemm + t_rot_slab * m
    del A, Am

    # --- Sigma_x as one GEMM over the sampled rows
    nx = cnt(8, q)
    Mo = rng.standard_normal((nx, o * naux))
    Mq = rng.standard_normal((min(q, 64), o * naux))
    t = _timeit(lambda: Mo @ Mq.T)
    stage["sigma_x"] = t * (q / nx) * (q / Mq.shape[0])
    desc["sigma_x"] = f"{nx}x{Mq.shape[0]} of {q}x{q} level pairs"
    del Mo, Mq

    # --- Sigma_c grid: one level, nW frequencies through the OpenMP kernel
    nW = cnt(2 * host_threads(), grid_steps)
    slabP = np.ascontiguousarray(slab.T)            # [P, m]
    e = np.sort(rng.uniform(-1, 3, n))
    pf = rng.uniform(0.3, 2.0, naux)
    pw = rng.uniform(0.1, 1.0, naux)
    om = np.linspace(-5, 5, nW)
    t = _timeit(lambda: sigma_ppm_diag(slabP, o, e, pf, pw, om))
    evals_per_level = grid_steps + 25               # grid + bisection / derivative evaluations
    stage["sigma_c"] = t / nW * evals_per_level * q
    desc["sigma_c"] = f"{nW} of {evals_per_level} frequency evaluations of 1 of {q} levels"

    # --- Sigma_c off-diagonal: weighted slabs (one kernel-cost pass per level) + GEMM q x q over (P, m)
    nl = cnt(4, q)
    Wl = rng.standard_normal((nl, n * naux // 8))
    t = _timeit(lambda: Wl @ Wl.T)
    stage["offdiag"] = t * 8 * (q / nl) ** 2 + (stage["sigma_c"] / (evals_per_level * q)) * q
    desc["offdiag"] = f"{nl}x{nl} of {q}x{q} pairs over 1/8 of the (P,m) range"
    del Wl

    # --- BSE setup: epsilon counted above; eigh + rotation of the whole tensor (reference rotates every slab)
    stage["bse_setup"] = t_eigh + t_rot_slab * m

    # --- Davidson: reference-structure matmul, H row block of one v1 = direct (vt x naux)(naux x ct^2)
    #     + exchange (ct x naux)(naux x vt ct), then row block times X
    k = 15
    nv = cnt(1, vt)
    Mcc = rng.standard_normal((naux, ct * ct))
    Mvc = rng.standard_normal((naux, vt * ct))
    Mv1v = rng.standard_normal((vt, naux))
    Mv1c = rng.standard_normal((ct, naux))
    X = rng.standard_normal((vt * ct, k))

    def row_block():
        for _ in range(nv):
            Hd = (Mv1v @ Mcc).reshape(vt, ct, ct)          # [v2, c1, c2]
            Hx = Mv1c @ Mvc                                # [c1, (v2 c2)]
            H = 2.0 * Hx - np.transpose(Hd, (1, 0, 2)).reshape(ct, vt * ct)
            H @ X
    t = _timeit(row_block) / nv
    n_solves = 2 if triplets else 1
    stage["davidson"] = t * vt * davidson_matmul_calls * n_solves
    desc["davidson"] = (f"{nv} of {vt} H row blocks of one matmul; {davidson_matmul_calls} matmul calls assumed per "
                        f"solve, {n_solves} solve(s)")
    total = float(sum(stage.values()))
    return {"seconds": total, "stage_seconds": {k2: round(v, 3) for k2, v in stage.items()}, "describe": desc,
            "threads": host_threads()}


def full_step(prob, nmax=10, grid_steps=1001, triplets=False):
    """Complete oracle step (small workloads only): wall seconds and results."""
    from . import gwbse_oracle as orc
    sz = prob["sizes"]
    gwopt = orc.GWOptions(sz.homo, sz.qpmin, sz.qpmax, sz.rpamin, sz.rpamax, qp_grid_steps=grid_steps)
    bseopt = orc.BSEOptions(sz.homo, sz.rpamin, sz.rpamax, sz.qpmin, sz.qpmax, sz.vmin, sz.cmax, nmax=nmax)
    t0 = time.perf_counter()
    res = orc.run_gwbse(prob["ao3c"], prob["C"], prob["energies"], prob["vxc"], prob["aux_coulomb"], gwopt, bseopt,
                        triplets=triplets)
    return time.perf_counter() - t0, res
