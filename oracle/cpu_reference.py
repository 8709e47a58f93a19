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


def _timeit(fn, reps=2):
    """Best of `reps` timed runs after one warm-up run (page faults, BLAS thread spin-up).  The short sleep lets the
    worker threads of the OTHER runtime (OpenBLAS after a GEMM sample, OpenMP after a C-kernel sample) stop spinning:
    without it a kernel sampled right after a GEMM was timed up to 2x too slow."""
    time.sleep(0.15)
    fn()
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def sampled_step(sz, davidson_matmul_calls=40, grid_steps=1001, triplets=False, scale=1.0, seed=7,
                 algorithm="reference", reps=2):
    """Estimated seconds per molecule for workload ``sz`` (xtp_b200.synth.Sizes) on the CPU, from a bounded sample of
    the independent units of every stage, scaled by the unit counts.  The sample sizes are fixed numbers (they do not
    depend on the core count); ``scale`` multiplies them (1.0 ~ 15-25 s of CPU work on 16 cores at C60 size).

    algorithm="reference":  the reference's loop structure (see the module docstring and oracle/cpu_step.py).
    algorithm="factorised": the CUDA path's algorithm on the CPU -- grid scan in one pass per slab, off-diagonal
        Sigma_c as a weighted-slab GEMM, no second eps(0)/eigensolver/rotation for the BSE, factorised BSE matmul.
    ``validate_against_full_step`` runs both this estimate and a complete step at a size where that is affordable."""
    assert algorithm in ("reference", "factorised")
    from xtp_b200 import synth
    use_all_host_threads()
    _base_timeit = globals()["_timeit"]

    def _timeit(fn, reps=reps):          # `reps` timed runs per sample (bench.py uses 1 when it takes the median of many steps)
        return _base_timeit(fn, reps)
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

    # --- aux rotation: per slab A^T (naux x naux) @ M[m] (naux x n)
    nr = cnt(2, m)
    A = rng.standard_normal((naux, naux))
    slab = rng.standard_normal((naux, n))          # [P, level]
    t_rot_slab = _timeit(lambda: [A.T @ slab for _ in range(nr)]) / nr
    desc["rotation"] = f"{nr} of {m} slabs"
    stage["metric"] = t_eigh + t_gemm + t_rot_slab * m      # one eigh + U s U^T + rotation
    # --- epsilon: per occupied level A_m diag(d) A_m^T
    nE = cnt(2, o)
    Am = np.ascontiguousarray(slab[:, :u])
    d = rng.random(u)
    t_eps_level = _timeit(lambda: [(Am * d) @ Am.T for _ in range(nE)]) / nE
    n_eps = 3 if algorithm == "reference" else 2
    desc["epsilon"] = (f"{nE} of {o} occupied levels per frequency; {n_eps} frequencies per step (PPM 2" +
                       (", BSE 1)" if algorithm == "reference" else "; the BSE reads eps(0) from the PPM eigenbasis)"))
    stage["epsilon"] = n_eps * t_eps_level * o
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

    # --- Sigma_c QP grid.  Energies as the synthetic workloads have them, plasmon-pole frequencies 0.3 .. 2 Ha.
    e = synth.make_energies(sz, np.random.default_rng(seed + 1))[sz.rpamin:sz.rpamax + 1]
    pf = rng.uniform(0.3, 2.0, naux)
    pw = rng.uniform(0.1, 1.0, naux)
    slab *= np.sqrt(synth.target_variance(sz))
    evals_per_level = grid_steps + 25               # grid + bisection / derivative evaluations
    if algorithm == "reference":
        nW = cnt(96, grid_steps)                    # frequencies of ONE level, one CalcCorrelationDiagElement each
        om = e[o - 1] + 0.01 * (np.arange(nW) - nW / 2)
        t = _timeit(lambda: sigma_ppm_diag(slab, o, e, pf, pw, om))
        stage["sigma_c"] = t / nW * evals_per_level * q
        desc["sigma_c"] = f"{nW} of {evals_per_level} frequency evaluations of 1 of {q} levels"
    else:
        from . import cpu_step
        lib = cpu_step._lib()
        nl = cnt(1, q)
        blk = np.ascontiguousarray(np.broadcast_to(slab, (nl,) + slab.shape))
        om0 = np.full(nl, e[o - 1] - 0.005 * (grid_steps - 1))
        vals = np.empty((nl, grid_steps))
        t = _timeit(lambda: lib.sigma_ppm_grid_batched(cpu_step._p(blk), naux * n, n, n, naux, o, cpu_step._p(e),
                                                       cpu_step._p(pf), cpu_step._p(pw), nl, cpu_step._p(om0), 0.01,
                                                       grid_steps, cpu_step._p(vals)), reps=1)
        om = e[o - 1] + 0.01 * np.arange(8)
        t1 = _timeit(lambda: sigma_ppm_diag(slab, o, e, pf, pw, om)) / 8
        stage["sigma_c"] = t / nl * q + t1 * 25 * q
        desc["sigma_c"] = (f"batched grid scan ({grid_steps} frequencies in one pass over the slab) of {nl} of {q} "
                           f"levels + 25 point evaluations per level")
        del blk

    # --- Sigma_c off-diagonal
    # throughput cost of one pass over one slab with every thread busy (the level-pair loop is spread over the threads)
    om_off = e[o - 1] + 0.05 * (np.arange(96) - 48)
    t_pass = _timeit(lambda: sigma_ppm_diag(slab, o, e, pf, pw, om_off)) / 96       # all threads busy: throughput per pass
    t_point = t_pass
    if algorithm == "reference":
        # CalcCorrelationOffDiagElement per level pair: two stabilised inverses per (P, m) = two point evaluations
        stage["offdiag"] = t_point * 2 * q * (q - 1) / 2
        desc["offdiag"] = f"2 slab passes per level pair, {q * (q - 1) // 2} pairs (pass cost from the Sigma_c sample)"
    else:
        nl = cnt(4, q)
        Wl = rng.standard_normal((nl, n * naux // 8))
        t = _timeit(lambda: Wl @ Wl.T)
        stage["offdiag"] = t * 8 * (q / nl) ** 2 + t_point * q
        desc["offdiag"] = f"weighted slabs + GEMM: {nl}x{nl} of {q}x{q} pairs over 1/8 of the (P,m) range"
        del Wl

    # --- BSE setup
    stage["bse_setup"] = (t_eigh + t_rot_slab * m) if algorithm == "reference" else 0.0

    # --- Davidson
    k = 15
    n_solves = 2 if triplets else 1
    X = rng.standard_normal((vt * ct, k))
    if algorithm == "reference":
        # matmul rebuilding H: row block of one v1 = direct (vt x naux)(naux x ct^2) + exchange (ct x naux)(naux x vt ct)
        nv = cnt(1, vt)
        Mcc = rng.standard_normal((naux, ct * ct))
        Mvc = rng.standard_normal((naux, vt * ct))
        Mv1v = rng.standard_normal((vt, naux))
        Mv1c = rng.standard_normal((ct, naux))

        def row_block():
            for _ in range(nv):
                Hd = (Mv1v @ Mcc).reshape(vt, ct, ct)          # [v2, c1, c2]
                Hx = Mv1c @ Mvc                                # [c1, (v2 c2)]
                H = 2.0 * Hx - np.transpose(Hd, (1, 0, 2)).reshape(ct, vt * ct)
                H @ X
        t = _timeit(row_block) / nv
        stage["davidson"] = t * vt * davidson_matmul_calls * n_solves
        desc["davidson"] = (f"{nv} of {vt} H row blocks of one matmul; {davidson_matmul_calls} matmul calls per solve, "
                            f"{n_solves} solve(s)")
    else:
        # factorised matmul on a 1/16 slice of the aux range (every term is a sum over P)
        pa = max(1, naux // 16)
        Mvc = rng.standard_normal((vt * ct, pa))
        Mcc = rng.standard_normal((pa * ct, ct))
        Mvv = rng.standard_normal((pa, vt, vt))
        Xt = rng.standard_normal((ct, vt * k))

        def fact():
            Mvc @ (Mvc.T @ X)
            U = (Mcc @ Xt).reshape(pa, ct, vt, k)
            np.einsum("pvw,pcwk->vck", Mvv, U, optimize=True)
        t = _timeit(fact) * (naux / pa)
        stage["davidson"] = t * davidson_matmul_calls * n_solves
        desc["davidson"] = (f"factorised matmul over {pa} of {naux} aux functions; {davidson_matmul_calls} matmul calls "
                            f"per solve, {n_solves} solve(s)")
    total = float(sum(stage.values()))
    return {"seconds": total, "stage_seconds": {k2: round(v, 3) for k2, v in stage.items()}, "describe": desc,
            "threads": host_threads(), "algorithm": algorithm}


def lowmem_problem(workload, seed=None):
    """synth.make_problem for shapes whose AO tensor is GBs (pentacene: 18 GB): slices generated chunk by chunk in
    place (same distribution as synth.make_ao3c)."""
    from xtp_b200 import synth
    sz = synth.WORKLOADS[workload] if isinstance(workload, str) else workload
    rng = np.random.default_rng(20260101 + sz.n_basis if seed is None else seed)
    nb = sz.n_basis
    prob = {"sizes": sz, "C": synth.make_mos(nb, rng), "energies": synth.make_energies(sz, rng)}
    dd = np.abs(np.arange(nb)[:, None] - np.arange(nb)[None, :])
    mask = np.exp(-dd / 32.0)
    mask *= np.sqrt(synth.target_variance(sz) / np.mean(mask * mask))
    ao = np.empty((sz.n_aux, nb, nb))
    for p0 in range(0, sz.n_aux, 64):
        p1 = min(sz.n_aux, p0 + 64)
        G = rng.standard_normal((p1 - p0, nb, nb))
        ao[p0:p1] = (G + np.transpose(G, (0, 2, 1))) * (mask / np.sqrt(2.0))
    prob["ao3c"] = ao
    prob["aux_coulomb"] = synth.make_aux_metric(sz, rng)
    prob["vxc"] = synth.make_vxc(sz, rng)
    return prob


def validate_against_full_step(workload="benzene-tzvp-shape", algorithm="reference"):
    """The sampled estimate next to a COMPLETE CPU step (oracle/cpu_step.py, nothing sampled or scaled) of the same
    shape and algorithm, stage by stage -- the extrapolation error as a number."""
    from . import cpu_step
    prob = lowmem_problem(workload)
    full = cpu_step.run_step(prob, algorithm=algorithm)
    est = sampled_step(prob["sizes"], davidson_matmul_calls=full["matmul_calls"], algorithm=algorithm)
    return {"shape": workload, "algorithm": algorithm, "cores": full["threads"],
            "full_s": round(full["seconds"], 3), "sampled_s": round(est["seconds"], 3),
            "sampled_over_full": round(est["seconds"] / full["seconds"], 3),
            "full_stage_seconds": full["stage_seconds"], "sampled_stage_seconds": est["stage_seconds"],
            "davidson_matmul_calls": full["matmul_calls"], "davidson_iterations": int(full["davidson_iterations"]),
            "qp_homo_ha": float(full["qp"][prob["sizes"].homo - prob["sizes"].qpmin]),
            "lowest_singlet_ha": float(full["singlets"][0])}


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
