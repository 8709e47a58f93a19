"""Worker of the multi-GPU parity test (launched by tests/test_gpu_multi.py under torchrun, one process per GPU).
Every rank drives its own libxtpb200 context; the library's NCCL communicator does the data-path collectives.
All host-visible results must equal the CPU oracle's on every rank (tolerances as in test_gpu_gwbse.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    return np.abs(a - b).max() / max(1e-300, np.abs(b).max())


def main():
    import torch

    from oracle import gwbse_oracle as orc
    from xtp_b200 import api, dist, synth

    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    rank, world, local = dist.init_process_group_from_env("nccl")
    ctx = api.Context(local)
    dist.join_library_communicator(ctx, rank, world)
    assert ctx.comm_info() == (rank, world)

    for name in ("tiny", "odd", "ch4-svp-shape"):
        if name == "odd":
            sz = synth.Sizes(n_basis=45, n_aux=91, homo=6, qpmax=15, cmax=15)
            prob = synth.make_problem(sz, seed=77)
        else:
            prob = synth.make_problem(name)
        sz = prob["sizes"]
        nmax = 3
        gwopt = orc.GWOptions(sz.homo, sz.qpmin, sz.qpmax, sz.rpamin, sz.rpamax, qp_grid_steps=201)
        bseopt = orc.BSEOptions(sz.homo, sz.rpamin, sz.rpamax, sz.qpmin, sz.qpmax, sz.vmin, sz.cmax, nmax=nmax,
                                davidson_tolerance="lapack")
        ref = orc.run_gwbse(prob["ao3c"], prob["C"], prob["energies"], prob["vxc"], prob["aux_coulomb"], gwopt, bseopt,
                            triplets=True)
        tc_o = orc.TCMatrix_gwbse().Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
        tc_o.Fill3cMO(prob["ao3c"], prob["C"])

        # -- set_raw / operator[] round trip through the cyclic column distribution
        tc = api.TCMatrix_gwbse(ctx).Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
        tc.set_raw(tc_o.M)
        np.testing.assert_array_equal(tc.get_raw(), tc_o.M)

        # -- collective Fill3cMO from this rank's packed AO slices (host buffer, then device-resident)
        p0, cnt = tc.local_aux_range()
        assert (p0, p0 + cnt) == dist.aux_range(sz.n_aux, rank, world)
        packed = api.pack_lower(prob["ao3c"][p0:p0 + cnt])
        pin = api.PinnedBuffer(max(1, packed.size))
        pin.array[:packed.size] = packed.reshape(-1)
        tc = api.TCMatrix_gwbse(ctx).Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
        tc.fill_begin(prob["C"])
        tc.fill_sharded_packed(packed_local=pin.array[:packed.size].reshape(packed.shape))
        assert rel(tc.get_raw(), tc_o.M) < 1e-12, name
        dev = torch.from_numpy(packed).cuda()
        tc.fill_begin(prob["C"])
        tc.fill_sharded_packed(dev_ptr=dev.data_ptr())
        assert rel(tc.get_raw(), tc_o.M) < 1e-12, name
        tc.apply_coulomb_metric(prob["aux_coulomb"])
        tc_o.MultiplyRightWithAuxMatrix(orc.Pseudo_InvSqrt_GWBSE(prob["aux_coulomb"], None)[0])
        assert rel(tc.get_raw(), tc_o.M) < 1e-10, name

        # -- RPA epsilon (partial sums over local unoccupied levels, all-reduced)
        rpa = api.RPA(tc)
        rpa.configure(sz.homo, sz.rpamin, sz.rpamax)
        rpa.setRPAInputEnergies(prob["energies"][sz.rpamin:sz.rpamax + 1])
        rpa_o = orc.RPA(tc_o)
        rpa_o.configure(sz.homo, sz.rpamin, sz.rpamax)
        rpa_o.setRPAInputEnergies(prob["energies"][sz.rpamin:sz.rpamax + 1])
        assert rel(rpa.calculate_epsilon_i(0.5), rpa_o.calculate_epsilon_i(0.5)) < 1e-12
        assert rel(rpa.calculate_epsilon_r(0.0), rpa_o.calculate_epsilon_r(0.0)) < 1e-12

        # -- GW (Sigma_x, PPM screening, QP grid solver, off-diagonal Sigma_c) and BSE singlets/triplets
        gw = api.GW(ctx, tc, prob["vxc"], prob["energies"])
        gw.configure(api.gw_options(homo=sz.homo, qpmin=sz.qpmin, qpmax=sz.qpmax, rpamin=sz.rpamin, rpamax=sz.rpamax,
                                    qp_grid_steps=201))
        gw.CalculateGWPerturbation()
        qp = gw.getGWAResults()
        np.testing.assert_allclose(qp, ref["qp_pert"], rtol=0, atol=1e-6)
        gw.CalculateHQP()
        hqp = gw.getHQP()
        np.testing.assert_allclose(hqp, ref["Hqp"], rtol=0, atol=1e-6)
        bse = api.BSE(ctx, tc)
        bse.configure(sz.homo, sz.rpamin, sz.rpamax, sz.qpmin, sz.qpmax, sz.vmin, sz.cmax, nmax,
                      gw.RPAInputEnergies(), hqp, davidson_tolerance="lapack")
        for dense_gb in ("32", "0"):      # H materialised (columns split over ranks) / factorised (aux split)
            os.environ["XTPB_BSE_DENSE_MAX_GB"] = dense_gb
            es, vs = bse.Solve_singlets_TDA()
            np.testing.assert_allclose(es, ref["singlet_energies"], rtol=0, atol=1e-6)
            et, _ = bse.Solve_triplets_TDA()
            np.testing.assert_allclose(et, ref["triplet_energies"], rtol=0, atol=1e-6)
        # identical on every rank (the host control flow depends on it)
        t = torch.from_numpy(np.concatenate([qp, es])).cuda()
        t0 = t.clone()
        torch.distributed.broadcast(t0, src=0)
        assert torch.equal(t, t0), "results differ between ranks"
        gw.close()
        bse.close()

        # -- Sigma_CDA with the quadrature nodes and the residue poles sharded over the ranks (BASELINE configs[2]):
        #    kernel values at arbitrary (level, frequency) pairs incl. frequencies that enclose poles, then a G0W0
        #    fixed-point solve, against the oracle to 1e-6 Ha
        if name != "odd":
            tco2 = orc.TCMatrix_gwbse().Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
            tco2.Fill(prob["ao3c"], prob["C"], prob["aux_coulomb"])
            tc2 = api.TCMatrix_gwbse(ctx).Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
            tc2.set_raw(tco2.M)
            kw = dict(sigma_integration="cda", qp_solver="fixedpoint", qp_grid_steps=41, order=12)
            gwc = api.GW(ctx, tc2, prob["vxc"], prob["energies"])
            gwc.configure(api.gw_options(homo=sz.homo, qpmin=sz.qpmin, qpmax=sz.qpmax, rpamin=sz.rpamin,
                                         rpamax=sz.rpamax, **kw))
            gwo = orc.GW(tco2, prob["vxc"], prob["energies"])
            gwo.configure(orc.GWOptions(sz.homo, sz.qpmin, sz.qpmax, sz.rpamin, sz.rpamax, **kw))
            gwc.PrepareScreening()
            gwo.rpa.setRPAInputEnergies(prob["energies"][sz.rpamin:sz.rpamax + 1])
            gwo.sigma.PrepareScreening()
            e = prob["energies"]
            rng = np.random.default_rng(4)
            levels = rng.integers(0, sz.qptotal, 10)
            freqs = np.concatenate([rng.uniform(-1.2, 1.2, 5), e[sz.qpmin + levels[5:]] + 0.05])
            val = gwc.CalcCorrelationDiagElements(levels, freqs)
            refv = np.array([gwo.sigma.CalcCorrelationDiagElement(int(l), float(x)) for l, x in zip(levels, freqs)])
            np.testing.assert_allclose(val, refv, rtol=1e-8, atol=1e-10)
            if name == "tiny":
                gwc.CalculateGWPerturbation()
                gwo.CalculateGWPerturbation()
                np.testing.assert_allclose(gwc.getGWAResults(), gwo.getGWAResults(), rtol=0, atol=1e-6)
            gwc.close()
        if rank == 0:
            print(f"mgpu ok: {name} world={world} S1={es[0]:.8f}", flush=True)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
