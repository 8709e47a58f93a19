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


def run(nb, ks, reps, out):
    homo = nb // 10 - 1
    sz = synth.Sizes(n_basis=nb, n_aux=3 * nb, homo=homo)
    rng = np.random.default_rng(20260101 + nb)
    ctx = api.Context(0)
    tc = api.TCMatrix_gwbse(ctx).Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
    tc.set_raw(synth.make_M_direct(sz, rng))
    hs = sz.vtotal + sz.ctotal
    hq = np.diag(np.sort(rng.uniform(-1.0, 2.0, hs)))
    eps_inv = rng.uniform(0.2, 1.0, sz.n_aux)
    vc, v, c, na = sz.bse_size, sz.vtotal, sz.ctotal, sz.n_aux
    for strategy in ("dense", "factorised"):
        os.environ["XTPB_BSE_DENSE_MAX_GB"] = "0" if strategy == "factorised" else "64"
        t0 = time.perf_counter()
        op = api.BSE_OPERATOR(ctx, 1, 2, 1, 0, eps_inv, tc, hq, sz.homo, sz.rpamin, sz.vmin, sz.cmax)
        ctx.sync()
        build_s = time.perf_counter() - t0
        for k in ks:
            X = np.linalg.qr(rng.standard_normal((vc, k)))[0]
            op.matmul(X)                                   # warm-up (includes the host<->device copies of X, Y)
            api.profile_reset()
            api.profile_enable(True)
            for _ in range(reps):
                op.matmul(X)
            api.profile_enable(False)
            p = api.profile_summary().get("bse_matmul", {"ms": 0.0, "work": 0.0, "launches": 0})
            ms = p["ms"] / reps
            flops = 4.0 * vc * na * k + 2.0 * na * k * vc * (v + c) + 2.0 * vc * k * (v + c)
            rec = {"n_basis": nb, "n_aux": na, "bse_size": vc, "strategy": strategy, "k": k,
                   "operator_build_s": round(build_s, 4), "kernel_ms_per_matmul": round(ms, 4),
                   "launches_per_matmul": p["launches"] / reps,
                   "algorithmic_tflops": round(flops / (ms * 1e-3) * 1e-12, 3) if ms > 0 else None,
                   "dense_stream_gbs": round((8.0 * vc * vc + 16.0 * vc * na) / (ms * 1e-3) * 1e-9, 1)
                   if ms > 0 and strategy == "dense" else None}
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
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        for nb in [int(x) for x in args.nb.split(",")]:
            run(nb, [int(x) for x in args.k.split(",")], args.reps, f)


if __name__ == "__main__":
    main()
