"""BASELINE.json configs[4]: synthetic scaling sweep of the BSE matvec -- random tensors, N_b 500..4000, N_aux = 3 N_b,
o = N_b/10 occupied levels, window v = c = o, k in {1, 10, 20, 40} trial vectors (SURVEY.md section 8d recipe).
Prints one JSON line per (N_b, strategy, k): seconds per BSE_OPERATOR::matmul call and the algorithmic rates.
Runs on the GPU box:
    python tools/bench_bse_matvec.py [--nb 500,1000,2000] [--reps 5] [--out gpurun_out/bse_matvec.jsonl]
The tensor is drawn directly in the device layout (synth.make_M_direct) for the m-window only; strategies:
  dense       screened direct term + Hqp materialised once, exchange factorised (the default when H fits)
  factorised  every term through the windows (what large problems fall back to)
Algorithmic flops per call: 4 vc N_aux k + 2 N_aux k vc (v + c) + 2 vc k (v + c); compulsory bytes of the
dense strategy: 8 (vc)^2 + 16 vc N_aux.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from xtp_b200 import api, synth  # noqa: E402

DMMA_PEAK = 37.09      # TFLOP/s, profiles/r01_fp64_probe.json
try:
    HBM_PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                 "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:  # noqa: BLE001
    HBM_PEAK = 6551.0


def run(nb, ks, reps, out, strategies):
    import torch
    homo = nb // 10 - 1
    # the operator needs the (v+c)^2 N_aux window of the tensor only: second index restricted to the BSE window
    sz = synth.Sizes(n_basis=nb, n_aux=3 * nb, homo=homo, rpamax=2 * homo + 1)
    rng = np.random.default_rng(20260101 + nb)
    ctx = api.Context(0)
    tc = api.TCMatrix_gwbse(ctx).Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
    Mdev = synth.draw_window_on_device(sz, 20260101 + nb)
    tc.set_raw_dev(Mdev.data_ptr())
    del Mdev
    torch.cuda.empty_cache()
    hs = sz.vtotal + sz.ctotal
    hq = np.diag(np.sort(rng.uniform(-1.0, 2.0, hs)))
    eps_inv = rng.uniform(0.2, 1.0, sz.n_aux)
    vc, v, c, na = sz.bse_size, sz.vtotal, sz.ctotal, sz.n_aux
    checks = {}
    for strategy in strategies:
        if strategy == "dense" and 8.0 * vc * vc > 64e9:
            continue                                       # H does not fit: the factorised operator is mandatory
        os.environ["XTPB_BSE_MODE"] = strategy
        os.environ["XTPB_BSE_DENSE_MAX_GB"] = "64"
        t0 = time.perf_counter()
        op = api.BSE_OPERATOR(ctx, 1, 2, 1, 0, eps_inv, tc, hq, sz.homo, sz.rpamin, sz.vmin, sz.cmax)
        ctx.sync()
        build_s = time.perf_counter() - t0
        for k in ks:
            X = np.linalg.qr(np.random.default_rng(k).standard_normal((vc, k)))[0]
            Y = op.matmul(X)                               # warm-up (includes the host<->device copies of X, Y)
            checks.setdefault(k, {})[strategy] = Y
            api.profile_reset()
            api.profile_enable(True)
            for _ in range(reps):
                op.matmul(X)
            api.profile_enable(False)
            p = api.profile_summary().get("bse_matmul", {"ms": 0.0, "work": 0.0, "launches": 0})
            ms = p["ms"] / reps
            flops = 4.0 * vc * na * k + 2.0 * na * k * vc * (v + c) + 2.0 * vc * k * (v + c)
            # compulsory HBM bytes per call: dense = H once + the flat exchange operand twice; factorised = the three
            # windows once -- plus X and Y
            nbytes = (8.0 * vc * vc + 16.0 * vc * na if strategy == "dense" else 8.0 * na * (vc + v * v + c * c)) \
                + 16.0 * vc * k
            tf = flops / (ms * 1e-3) * 1e-12 if ms > 0 else 0.0
            gbs = nbytes / (ms * 1e-3) * 1e-9 if ms > 0 else 0.0
            both = checks[k]
            rec = {"n_basis": nb, "n_aux": na, "bse_size": vc, "strategy": strategy, "k": k,
                   "operator_build_s": round(build_s, 4), "kernel_ms_per_matmul": round(ms, 4),
                   "launches_per_matmul": p["launches"] / reps,
                   "algorithmic_tflops": round(tf, 3), "dmma_frac": round(tf / DMMA_PEAK, 4),
                   "compulsory_gbs": round(gbs, 1), "hbm_frac": round(gbs / HBM_PEAK, 4),
                   "bound": "tensor" if tf / DMMA_PEAK > gbs / HBM_PEAK else "hbm",
                   "dense_vs_factorised_rel_diff": (float(np.abs(both["dense"] - both["factorised"]).max()
                                                          / np.abs(both["dense"]).max())
                                                    if len(both) == 2 else None)}
            line = json.dumps(rec)
            print(line, flush=True)
            out.write(line + "\n")
        op.close()
    tc.close()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nb", default="500,1000,2000")
    ap.add_argument("--k", default="1,10,20,40")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="gpurun_out/bse_matvec.jsonl")
    ap.add_argument("--strategies", default="dense,factorised")
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        for nb in [int(x) for x in args.nb.split(",")]:
            run(nb, [int(x) for x in args.k.split(",")], args.reps, f, args.strategies.split(","))


if __name__ == "__main__":
    main()
