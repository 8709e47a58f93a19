"""world_size-2 `gloo` tests (CPU) of the multi-GPU host logic (xtp_b200/dist.py).

The CUDA library distributes the tensor over its second index (cyclic) and the BSE operator over the aux index and
all-reduces partial results.  Here two CPU processes replay exactly that partition with the numpy oracle and a gloo
all-reduce / all-gather, and every rank must recover the unsharded oracle result -- this pins the index arithmetic
(local column maps, local occupied counts, canonical aux ranges, the Fill3cMO gather schedule)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from oracle import gwbse_oracle as orc
from xtp_b200 import dist, synth


def _allreduce(a):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).copy())
    tdist.all_reduce(t)
    return t.numpy()


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = dist.init_process_group_from_env("gloo")
    assert (r, w) == (rank, world)
    try:
        sz = synth.Sizes(n_basis=23, n_aux=37, homo=4, qpmax=11, cmax=11)      # odd everywhere, ragged splits
        prob = synth.make_problem(sz, seed=5)
        e = prob["energies"][sz.rpamin:sz.rpamax + 1]
        tc = orc.TCMatrix_gwbse().Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
        tc.Fill3cMO(prob["ao3c"], prob["C"])
        M = tc.M                                                    # [m, P, n]
        cols = dist.local_columns(sz.ntotal, rank, world)
        n_occ = sz.n_occ
        n_occ_loc = dist.n_local_below(n_occ, rank, world)
        assert n_occ_loc == int((cols < n_occ).sum())
        Ml, el = M[:, :, cols], e[cols]

        # ---- Fill3cMO schedule: first half split over the aux index, all-gather, second half over local columns
        lo, hi = dist.aux_range(sz.n_aux, rank, world)
        Cm = prob["C"][:, sz.rpamin:sz.mmax + 1]
        Cn = prob["C"][:, sz.rpamin + cols]
        W_loc = np.einsum('pab,bm->pam', prob["ao3c"][lo:hi], Cm)           # T_P C_m for the local aux range
        maxcnt = max(dist.aux_range(sz.n_aux, s, world)[1] - dist.aux_range(sz.n_aux, s, world)[0] for s in range(world))
        pad = np.zeros((maxcnt,) + W_loc.shape[1:])
        pad[:hi - lo] = W_loc
        gathered = [torch.zeros(pad.shape, dtype=torch.float64) for _ in range(world)]
        tdist.all_gather(gathered, torch.from_numpy(pad))
        M_fill = np.zeros((sz.mtotal, sz.n_aux, len(cols)))
        for s in range(world):
            a, b = dist.aux_range(sz.n_aux, s, world)
            M_fill[:, a:b, :] = np.einsum('an,pam->mpn', Cn, gathered[s].numpy()[:b - a])
        np.testing.assert_allclose(M_fill, Ml, rtol=0, atol=1e-13)

        # ---- epsilon: all occupied m (first index), local unoccupied a (second index), all-reduce
        rpa = orc.RPA(tc)
        rpa.configure(sz.homo, sz.rpamin, sz.rpamax)
        rpa.setRPAInputEnergies(e)
        for omega, imag in ((0.5, True), (0.0, False)):
            part = np.zeros((sz.n_aux, sz.n_aux))
            dE = el[n_occ_loc:][None, :] - e[:n_occ][:, None]
            if imag:
                d = 4.0 * dE / (dE * dE + omega * omega)
            else:
                d = 2.0 * ((dE - omega) / ((dE - omega) ** 2 + rpa.eta ** 2) + (dE + omega) / ((dE + omega) ** 2 + rpa.eta ** 2))
            for m in range(n_occ):
                A = Ml[m][:, n_occ_loc:]
                part += (A * d[m][None, :]) @ A.T
            eps = _allreduce(part) + np.eye(sz.n_aux)
            full = rpa.calculate_epsilon_i(omega) if imag else rpa.calculate_epsilon_r(omega)
            np.testing.assert_allclose(eps, full, rtol=0, atol=1e-12)

        # ---- Sigma_x: partial sums over the local occupied second-index levels
        gwopt = orc.GWOptions(sz.homo, sz.qpmin, sz.qpmax, sz.rpamin, sz.rpamax, qp_grid_steps=51)
        sig = orc.Sigma_PPM(tc, rpa)
        sig.configure(gwopt)
        sx_full = sig.CalcExchangeMatrix()
        q0, q = sz.qpmin - sz.rpamin, sz.qptotal
        occ = Ml[q0:q0 + q][:, :, :n_occ_loc]
        sx = _allreduce(-np.einsum('npm,kpm->nk', occ, occ))
        np.testing.assert_allclose(sx, sx_full, rtol=0, atol=1e-12)

        # ---- Sigma_c (PPM) diagonal: partial pole sums over the local columns with the local occupied count
        sig.PrepareScreening()                        # rotates tc.M in place (identically on both ranks)
        Ml = tc.M[:, :, cols]
        fac, Om = sig._fac(), sig.ppm.ppm_freq
        for level, freq in ((0, -0.6), (q - 1, 0.31)):
            x = freq - el[None, :] + np.zeros((sz.n_aux, 1))
            x[:, :n_occ_loc] += Om[:, None]
            x[:, n_occ_loc:] -= Om[:, None]
            g = orc.ppm_stabilized_inverse(x)
            slab = Ml[q0 + level]
            val = _allreduce(np.array([(fac[:, None] * g * slab * slab).sum()]))[0]
            assert abs(val - sig.CalcCorrelationDiagElement(level, freq)) < 1e-12

        # ---- BSE operator distributed over the aux index: Y = sum over ranks of the partial products
        v0, c0 = sz.vmin - sz.rpamin, sz.homo + 1 - sz.rpamin
        vt, ct = sz.vtotal, sz.ctotal
        rng = np.random.default_rng(11)
        X = rng.standard_normal((vt * ct, 3))
        eps_inv = rng.uniform(0.3, 1.0, sz.n_aux)
        P = slice(*dist.aux_range(sz.n_aux, rank, world))
        Mf = tc.M
        Mvc = Mf[v0:v0 + vt, P, c0:c0 + ct]
        Mvv = Mf[v0:v0 + vt, P, v0:v0 + vt]
        Mcc = Mf[c0:c0 + ct, P, c0:c0 + ct]
        Xr = X.reshape(vt, ct, -1)
        T = np.einsum('vpc,vck->pk', Mvc, Xr)
        Yx = np.einsum('vpc,pk->vck', Mvc, T)
        Yd = np.einsum('vpw,p,cpd,wdk->vck', Mvv, eps_inv[P], Mcc, Xr)
        Y = _allreduce((2.0 * Yx - Yd).reshape(vt * ct, -1))
        Mvc_f, Mvv_f, Mcc_f = Mf[v0:v0 + vt, :, c0:c0 + ct], Mf[v0:v0 + vt, :, v0:v0 + vt], Mf[c0:c0 + ct, :, c0:c0 + ct]
        Hx = np.einsum('vpc,wpd->vcwd', Mvc_f, Mvc_f).reshape(vt * ct, vt * ct)
        Hd = np.einsum('vpw,p,cpd->vcwd', Mvv_f, eps_inv, Mcc_f).reshape(vt * ct, vt * ct)
        np.testing.assert_allclose(Y, (2.0 * Hx - Hd) @ X, rtol=0, atol=1e-11)

        # ---- BSE operator, dense mode (bse.cu: BseOperator with H in HBM): rank r owns the columns (v2, c2) with v2 in
        # [vt r / world, vt (r+1) / world); the screened direct term is the dense block H[:, owned], the exchange term
        # stays factorised with the PARTIAL T of the owned rows (linear in T, so the all-reduce of Y completes it)
        v2lo, v2hi = vt * rank // world, vt * (rank + 1) // world
        own = slice(v2lo * ct, v2hi * ct)
        Fvc = np.transpose(Mvc_f, (1, 0, 2)).reshape(sz.n_aux, vt * ct)           # flat operand [P][(v, c)]
        T_part = Fvc[:, own] @ X[own]
        Y_dense = _allreduce(-Hd[:, own] @ X[own] + 2.0 * Fvc.T @ T_part)
        np.testing.assert_allclose(Y_dense, (2.0 * Hx - Hd) @ X, rtol=0, atol=1e-11)
        d_part = np.zeros(vt * ct)
        d_part[own] = -np.diag(Hd)[own] + 2.0 * (Fvc[:, own] ** 2).sum(axis=0)
        np.testing.assert_allclose(_allreduce(d_part), np.diag(2.0 * Hx - Hd), rtol=0, atol=1e-12)

        # ---- compressed QP-grid scan: every rank plans over ITS columns (own pole range, own bins) and the partial
        # grids are summed; host plan from libxtpb200 (no device), kernels replayed by the numpy mirror
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import ppm_grid_mirror as mir
        steps, spacing, level = 257, 0.02, q0 + 1
        g0 = np.array([e[sz.qpmin - sz.rpamin + 1] - spacing * (steps - 1) / 2])
        zmin, zmax = mir.pole_range(el, n_occ_loc, Om, fac)
        pl = mir.plan(g0, spacing, steps, zmin, zmax)
        assert pl is not None
        part = mir.grid_values(np.ascontiguousarray(Ml[level]), el, n_occ_loc, Om, fac, g0[0], spacing, steps, pl[0], pl[1][0])
        full = mir.direct_values(tc.M[level], e, n_occ, Om, fac, g0[0], spacing, steps)
        np.testing.assert_allclose(_allreduce(part), full, rtol=1e-10, atol=1e-12)
        # ---- collective Fill3cMO, second half in ONE batched product per round: the gathered half-transformed blocks
        # sit in slots (source rank, index in the round); slot -> aux index as k_scatter_fill_slots maps it
        B = 5
        rounds = (maxcnt + B - 1) // B
        M_slots = np.zeros((sz.mtotal, sz.n_aux, len(cols)))
        for i in range(rounds):
            for s_ in range(world):
                a, b = dist.aux_range(sz.n_aux, s_, world)
                for k_ in range(B):
                    P_ = a + i * B + k_
                    if P_ < b:                                      # other slots are padding of the last round
                        M_slots[:, P_, :] = np.einsum('an,am->mn', Cn, gathered[s_].numpy()[i * B + k_])
        np.testing.assert_allclose(M_slots, M[:, :, cols], rtol=0, atol=1e-13)

        # ---- Sigma_CDA with the quadrature nodes and the residue poles sharded over the ranks (sigma_other.cu):
        # eps(i w_j) is summed onto rank j % world only (reduce), inverted there, and broadcast; every rank forms the
        # quadratic forms over ITS columns; a residue's eps(|e_i - w|) is summed onto rank k % world, which solves
        # eps x = v with the vector v gathered from the rank that owns second-index column i
        tc2 = orc.TCMatrix_gwbse().Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
        tc2.Fill(prob["ao3c"], prob["C"], prob["aux_coulomb"])
        rpa2 = orc.RPA(tc2)
        rpa2.configure(sz.homo, sz.rpamin, sz.rpamax)
        rpa2.setRPAInputEnergies(e)
        cda = orc.Sigma_CDA(tc2, rpa2)
        cda.configure(orc.SigmaOptions(sz.homo, sz.qpmin, sz.qpmax, sz.rpamin, sz.rpamax, order=8))
        cda.PrepareScreening()
        M2l = tc2.M[:, :, cols]

        def eps_partial(omega, imag):
            dE = el[n_occ_loc:][None, :] - e[:n_occ][:, None]
            if imag:
                d = 4.0 * dE / (dE * dE + omega * omega)
            else:
                d = 2.0 * ((dE - omega) / ((dE - omega) ** 2 + rpa2.eta ** 2) + (dE + omega) / ((dE + omega) ** 2 + rpa2.eta ** 2))
            part = np.zeros((sz.n_aux, sz.n_aux))
            for m in range(n_occ):
                A = M2l[m][:, n_occ_loc:]
                part += (A * d[m][None, :]) @ A.T
            return part

        def reduce_to(part, root):
            t = torch.from_numpy(part.copy())
            tdist.reduce(t, dst=root)
            return t.numpy()

        def bcast(a, root):
            t = torch.from_numpy(np.ascontiguousarray(a).copy())
            tdist.broadcast(t, src=root)
            return t.numpy()

        order = cda.gq.Order()
        eye = np.eye(sz.n_aux)
        k0 = reduce_to(eps_partial(0.0, False), order % world)
        kappa0 = bcast(np.linalg.inv(k0 + eye) - eye if rank == order % world else np.zeros_like(k0), order % world)
        np.testing.assert_allclose(kappa0, cda.kappa0, rtol=0, atol=1e-10)
        kernels = []
        for j in range(order):
            wj = cda.gq.ScaledPoint(j)
            full = reduce_to(eps_partial(wj, True), j % world)
            mine = -(np.linalg.inv(full + eye) - eye) + np.exp(-(cda.opt.alpha * wj) ** 2) * kappa0 \
                if rank == j % world else np.zeros_like(full)
            kernels.append(bcast(mine, j % world))
            np.testing.assert_allclose(kernels[j], cda.dielinv[j], rtol=0, atol=1e-10)
        level, freq = 2, e[q0 + 2] + 0.07
        slab_l = M2l[q0 + level]                                      # [P, local m]
        dEc = (freq - el).astype(np.complex128)
        dEc[:n_occ_loc] += 1j * rpa2.getEta()
        dEc[n_occ_loc:] -= 1j * rpa2.getEta()
        gq = 0.0
        for j in range(order):
            wj = cda.gq.ScaledPoint(j)
            den = (1.0 / (dEc + 1j * wj) + 1.0 / (dEc - 1j * wj)).real
            gq += cda.gq.ScaledWeight(j) * float((den * np.einsum('pm,pm->m', slab_l, kernels[j] @ slab_l)).sum())
        gq = _allreduce(np.array([0.5 / np.pi * gq]))[0]
        np.testing.assert_allclose(gq, cda.SigmaGQDiag(freq, level), rtol=0, atol=1e-11)
        # residues: items = enclosed poles (identical list on every rank), item k evaluated by rank k % world
        fermi = 0.5 * (e[n_occ - 1] + e[n_occ])
        items = [(i, cda.CalcResiduePrefactor(fermi, e[i], freq)) for i in range(len(e))]
        items = [(i, f_) for i, f_ in items if abs(f_) > 1e-10]
        assert items, "pick a frequency that encloses at least one pole"
        res = 0.0
        for k_, (i, f_) in enumerate(items):
            full = reduce_to(eps_partial(abs(e[i] - freq), False), k_ % world)
            v = np.zeros(sz.n_aux)
            if i % world == rank:                                     # owner of second-index column i
                v = tc2.M[q0 + level][:, i].copy()
            v = _allreduce(v)
            if k_ % world == rank:
                x = np.linalg.solve(full + eye, v)
                res += f_ * float(v @ x - v @ v)
        res = _allreduce(np.array([res]))[0]
        tail_free = orc.Sigma_CDA(tc2, rpa2)
        tail_free.configure(orc.SigmaOptions(sz.homo, sz.qpmin, sz.qpmax, sz.rpamin, sz.rpamax, order=8, alpha=0.0))
        tail_free.kappa0, tail_free.dielinv, tail_free.gq = cda.kappa0, cda.dielinv, cda.gq
        np.testing.assert_allclose(res, tail_free.CalcResidueContribution(freq, level), rtol=0, atol=1e-10)
        # ---- replicated N_aux^3 products split by column blocks (tc.cu: congruence_sym and the folded rotation):
        # every rank holds E (symmetric) and a general R; rank r forms columns [n r / world, n (r+1) / world) of E R and
        # of R^T (E R); one broadcast per block completes the matrix on every rank
        rng2 = np.random.default_rng(77)
        na = sz.n_aux
        Es = rng2.standard_normal((na, na))
        Es = Es + Es.T
        Rg = np.triu(rng2.standard_normal((na, na))) + 3.0 * np.eye(na)        # e.g. the Cholesky factor of the metric
        a, b = dist.aux_range(na, rank, world)
        Tblk = Es @ Rg[:, a:b]
        out = np.zeros((na, na))
        out[:, a:b] = Rg.T @ Tblk
        for r_ in range(world):
            a_, b_ = dist.aux_range(na, r_, world)
            blk = torch.from_numpy(np.ascontiguousarray(out[:, a_:b_]))
            tdist.broadcast(blk, src=r_)
            out[:, a_:b_] = blk.numpy()
        np.testing.assert_allclose(out, Rg.T @ Es @ Rg, rtol=1e-12, atol=1e-10)
        results[rank] = "ok"
    except Exception as exc:  # noqa: BLE001
        import traceback
        results[rank] = traceback.format_exc() + str(exc)
    finally:
        tdist.barrier()
        tdist.destroy_process_group()


def test_partition_helpers():
    for n, world in ((23, 2), (7, 3), (1860, 8), (5, 5)):
        seen = np.concatenate([dist.local_columns(n, r, world) for r in range(world)])
        assert sorted(seen) == list(range(n))
        for g in (0, 1, n // 2, n):
            assert sum(dist.n_local_below(g, r, world) for r in range(world)) == g
        edges = [dist.aux_range(n, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == n and all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))


@pytest.mark.timeout(300)
def test_sharded_pipeline_world2_gloo():
    world = 2
    port = 29500 + os.getpid() % 400
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
        assert dict(results) == {0: "ok", 1: "ok"}, dict(results)
