"""Times the fused Sigma_c PPM grid kernel (GW::SolveQP_Grid scan) on a synthetic tensor.
    python tools/bench_sigma_grid.py [--workload synth-1000] [--out gpurun_out/sigma_grid.jsonl]"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(workload, reps):
    import numpy as np

    from xtp_b200 import api, synth
    sz = synth.WORKLOADS[workload]
    rng = np.random.default_rng(3)
    ctx = api.Context(0)
    tc = api.TCMatrix_gwbse(ctx).Initialize(sz.n_aux, sz.rpamin, sz.mmax, sz.rpamin, sz.rpamax)
    tc.set_raw(synth.make_M_direct(sz, rng))
    e = synth.make_energies(sz, rng)
    gw = api.GW(ctx, tc, synth.make_vxc(sz, rng), e)
    gw.configure(api.gw_options(homo=sz.homo, qpmin=sz.qpmin, qpmax=sz.qpmax, rpamin=sz.rpamin, rpamax=sz.rpamax))
    gw.PrepareScreening()
    centers = e[sz.qpmin:sz.qpmax + 1].copy()
    out = gw.CalcCorrelationGrid(centers)          # warm-up
    api.profile_reset()
    api.profile_enable(True)
    for _ in range(reps):
        out = gw.CalcCorrelationGrid(centers)
    api.profile_enable(False)
    p = api.profile_summary()["sigma_ppm_grid"]
    info = gw.grid_scan_info()
    print(json.dumps({"workload": workload, "compressed": info["compressed"], "bins": info["bins"],
                      "evaluated_fraction": info["direct_evaluations"] / info["equivalent_evaluations"],

                      "ms": p["ms"] / reps, "gevals_per_s": p["work"] / (p["ms"] * 1e-3) * 1e-9,
                      "checksum": float(np.abs(out).sum()), "sample": out[1, 498:503].tolist()}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="synth-1000")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--out", default="gpurun_out/sigma_grid.jsonl")
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--walk-only", action="store_true", help="only the three near-range walks (XTPB_GRID_WALK)")
    args = ap.parse_args()
    if args.child:
        child(args.workload, args.reps)
        return
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        # the two ways the scan can run: pole by pole and compressed; for the compressed scan the three walks of the
        # near m-ranges (XTPB_GRID_WALK) and register/occupancy caps around the default
        variants = [("direct", {}), ("compressed", {"XTPB_GRID_KERNEL": "warp", "XTPB_GRID_WALK": "0"}),
                    ("compressed", {"XTPB_GRID_KERNEL": "warp", "XTPB_GRID_WALK": "1"}),
                    ("compressed", {"XTPB_GRID_KERNEL": "warp", "XTPB_GRID_WALK": "2"}),
                    ("compressed", {"XTPB_GRID_KERNEL": "cta", "XTPB_GRID_OCC": "3"}),
                    ("compressed", {"XTPB_GRID_KERNEL": "cta", "XTPB_GRID_OCC": "4"}),
                    ("compressed", {"XTPB_GRID_KERNEL": "cta", "XTPB_GRID_OCC": "5"})]
        if args.walk_only:
            variants = [v for v in variants if v[1].get("XTPB_GRID_WALK") == "2" or v[1].get("XTPB_GRID_KERNEL") == "cta"]
        for mode, extra in variants:
            env = dict(os.environ, XTPB_SIGMA_GRID=mode)
            env.update(extra)
            r = subprocess.run([sys.executable, __file__, "--child", "--workload", args.workload, "--reps",
                                str(args.reps)], env=env, capture_output=True, text=True)
            if r.stdout.strip():
                rec = json.loads(r.stdout.strip().splitlines()[-1])
            else:
                rec = {"error": r.stderr[-400:]}
            rec.update({"mode": mode, "env": extra})
            line = json.dumps(rec)
            print(line, flush=True)
            f.write(line + "\n")


if __name__ == "__main__":
    main()
