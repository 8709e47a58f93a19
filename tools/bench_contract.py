"""Times the DMMA contraction engine on the shapes the GW-BSE path launches (C60-tzvp-shape sizes unless --scale)
and prints achieved FP64 TFLOP/s per shape as JSON lines.  Runs on the GPU box:
    python tools/bench_contract.py [--reps 5] [--out gpurun_out/contract_sweep.jsonl]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from xtp_b200 import _lib, api  # noqa: E402


def desc(**kw):
    d = _lib.ContractDesc()
    d.n_outer = 1
    d.n_batch = 1
    d.alpha = 1.0
    d.beta = 0.0
    d.force_cfg = -1
    d.force_splits = 0
    for k, v in kw.items():
        setattr(d, k, v)
    return d


def shapes(nb, naux, homo):
    o = homo + 1
    n = nb
    m = 2 * o
    u = n - o
    v = c = o
    ld = n + (n & 1)
    out = []
    # eps(w): SYRK lower, K-contig both, n_outer = occupied levels, weights
    a0 = o & ~1
    K = n - a0
    out.append(("epsilon_syrk", desc(M=naux, N=naux, K=K, n_outer=o, lower=1,
                                     a_row=ld, a_k=1, a_outer=naux * ld, a_len=o * naux * ld,
                                     b_row=ld, b_k=1, b_outer=naux * ld, b_len=o * naux * ld,
                                     c_row=1, c_col=naux, c_len=naux * naux,
                                     d_outer=K, d_len=o * K), float(naux) * naux * K * o))
    # K2 rotation: per slab (n x naux)(naux x naux); A rows-contig, B K-contig; batch 32 slabs
    nbat = 32
    out.append(("aux_rotation", desc(M=n, N=naux, K=naux, n_batch=nbat,
                                     a_row=1, a_k=ld, a_batch=naux * ld, a_len=nbat * naux * ld,
                                     b_row=naux, b_k=1, b_len=naux * naux,
                                     c_row=1, c_col=ld, c_batch=naux * ld, c_len=nbat * naux * ld),
                2.0 * n * naux * naux * nbat))
    # K1 step 1: W_P = T_P C_m : M=nb, N=m, K=nb, batch 32 aux functions
    ldt = nb + (nb & 1)
    out.append(("fill_T_times_Cm", desc(M=nb, N=m, K=nb, n_batch=nbat,
                                        a_row=ldt, a_k=1, a_batch=ldt * nb, a_len=nbat * ldt * nb,
                                        b_row=ldt, b_k=1, b_len=ldt * m,
                                        c_row=1, c_col=ldt, c_batch=ldt * m, c_len=nbat * ldt * m),
                2.0 * nb * m * nb * nbat))
    # K1 step 2: M[m][P][:] = C_n^T W_P : M=n, N=m, K=nb, batch 32; output scattered into [m][P][n]
    out.append(("fill_CnT_times_W", desc(M=n, N=m, K=nb, n_batch=nbat,
                                         a_row=ldt, a_k=1, a_len=ldt * n,
                                         b_row=ldt, b_k=1, b_batch=ldt * m, b_len=nbat * ldt * m,
                                         c_row=1, c_col=naux * ld, c_batch=ld, c_len=m * naux * ld),
                2.0 * n * m * nb * nbat))
    # BSE direct step 1: rows (c1,P) x cols (k,v2), K = ct
    k = 20
    ldc = c + (c & 1)
    ldu = v + (v & 1)
    out.append(("bse_direct_step1", desc(M=c * naux, N=k * v, K=c,
                                         a_row=ldc, a_k=1, a_len=c * naux * ldc,
                                         b_row=c, b_k=1, b_len=k * v * c,
                                         c_row=ldu, c_col=1, c_col_inner=v, c_col_outer=c * naux * ldu,
                                         c_len=k * c * naux * ldu), 2.0 * c * naux * k * v * c))
    # BSE direct step 2: rows (k,c1) x cols v1, K = v, n_outer = naux
    out.append(("bse_direct_step2", desc(M=k * c, N=v, K=v, n_outer=naux,
                                         a_row=naux * ldu, a_k=1, a_outer=ldu, a_len=k * c * naux * ldu,
                                         b_row=naux * ldu, b_k=1, b_outer=ldu, b_len=v * naux * ldu,
                                         c_row=1, c_col=k * c, c_len=k * c * v), 2.0 * k * c * v * v * naux))
    # BSE exchange: T = Mvc^T X  (M=naux, N=k, K=ct, n_outer=vt)
    out.append(("bse_exchange_T", desc(M=naux, N=k, K=c, n_outer=v,
                                       a_row=ldc, a_k=1, a_outer=naux * ldc, a_len=v * naux * ldc,
                                       b_row=v * c, b_k=1, b_outer=c, b_len=k * v * c,
                                       c_row=1, c_col=naux, c_len=naux * k), 2.0 * naux * k * c * v))
    # Sigma_x: q x q, K = n_occ, n_outer = naux
    q = m
    out.append(("sigma_x", desc(M=q, N=q, K=o, n_outer=naux, lower=1,
                                a_row=naux * ld, a_k=1, a_outer=ld, a_len=q * naux * ld,
                                b_row=naux * ld, b_k=1, b_outer=ld, b_len=q * naux * ld,
                                c_row=1, c_col=q, c_len=q * q), float(q) * q * o * naux))
    # dense BSE direct term over the occupied pairs (bse.cu): rows (c2,c1), a slice of the pair columns, K = naux;
    # both operands row-contiguous with long k strides
    npair = v * (v + 1) // 2
    ncol = min(npair, 8192)
    lda, ldb = c * c + (c * c & 1), npair + (npair & 1)
    out.append(("bse_dense_direct_pairs", desc(M=c * c, N=ncol, K=naux,
                                               a_row=1, a_k=lda, a_len=naux * lda,
                                               b_row=1, b_k=ldb, b_len=naux * ldb,
                                               c_row=1, c_col=c * c, c_len=c * c * ncol), 2.0 * c * c * ncol * naux))
    # plain square DGEMM for reference
    for s in (4096, 8192):
        out.append((f"square_{s}_kc_kc", desc(M=s, N=s, K=s, a_row=s, a_k=1, a_len=s * s, b_row=s, b_k=1,
                                              b_len=s * s, c_row=1, c_col=s, c_len=s * s), 2.0 * s ** 3))
        out.append((f"square_{s}_mc_kc", desc(M=s, N=s, K=s, a_row=1, a_k=s, a_len=s * s, b_row=s, b_k=1,
                                              b_len=s * s, c_row=1, c_col=s, c_len=s * s), 2.0 * s ** 3))
        out.append((f"square_{s}_mc_mc", desc(M=s, N=s, K=s, a_row=1, a_k=s, a_len=s * s, b_row=1, b_k=s,
                                              b_len=s * s, c_row=1, c_col=s, c_len=s * s), 2.0 * s ** 3))
    return out


def cublas_reference(nb, naux, homo, reps):
    """What the reference's GPU path runs for the same stages: cuBLAS through torch's FP64 matmul (cublasDgemm /
    cublasDgemmStridedBatched) on the same shapes, same box.  epsilon follows OpenMP_CUDA::A_TDA (upstream
    openmp_cuda.cc / cudapipeline.cc): per occupied level a row scaling (Ddgmm) and a full Dgemm accumulated into the
    N_aux^2 result -- 2 o u N_aux^2 flops where the SYRK-shaped launch of this library needs half."""
    import torch
    dev = torch.device("cuda", 0)
    o = homo + 1
    n, m, u = nb, 2 * o, nb - o
    g = torch.Generator(device=dev)
    g.manual_seed(1)

    def rnd(*shape):
        return torch.randn(shape, dtype=torch.float64, device=dev, generator=g) * 1e-2

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    out = {}
    nbat = 32
    # aux rotation: out[m] = R^T M[m], M[m] is (N_aux x n) row-major = the library's [P][n] slab
    R, Mb = rnd(naux, naux), rnd(nbat, naux, n)
    Ob = torch.empty_like(Mb)
    ms = timed(lambda: torch.matmul(R.T, Mb, out=Ob))
    out["aux_rotation"] = (ms, 2.0 * n * naux * naux * nbat)
    del R, Mb, Ob
    # Fill3cMO: W_P = T_P C_m, then C_n^T W_P, batched over 32 aux functions
    Tb, Cm, Cn = rnd(nbat, nb, nb), rnd(nb, m), rnd(nb, n)
    W = torch.empty((nbat, nb, m), dtype=torch.float64, device=dev)
    ms = timed(lambda: torch.matmul(Tb, Cm, out=W))
    out["fill_T_times_Cm"] = (ms, 2.0 * nb * m * nb * nbat)
    O2 = torch.empty((nbat, n, m), dtype=torch.float64, device=dev)
    ms = timed(lambda: torch.matmul(Cn.T, W, out=O2))
    out["fill_CnT_times_W"] = (ms, 2.0 * n * m * nb * nbat)
    del Tb, W, O2
    # epsilon: per occupied level A (N_aux x u): acc += (A * d) A^T   (Ddgmm + Dgemm), a sample of 8 levels scaled to o
    A, d = rnd(8, naux, u), torch.rand((8, u), dtype=torch.float64, device=dev, generator=g)
    acc = torch.zeros((naux, naux), dtype=torch.float64, device=dev)

    def eps():
        for i in range(8):
            acc.addmm_(A[i] * d[i], A[i].T)
    ms = timed(eps) * o / 8.0
    K = u
    out["epsilon_syrk"] = (ms, float(naux) * naux * K * o)           # algorithmic (SYRK-minimal) flops, like the library line
    out["epsilon_gemm_equivalent_tflops"] = 2.0 * naux * naux * K * o / ms * 1e-9
    del A, acc
    for s_ in (4096, 8192):
        X, Y = rnd(s_, s_), rnd(s_, s_)
        Z = torch.empty_like(X)
        ms = timed(lambda: torch.matmul(X, Y, out=Z))
        out[f"square_{s_}_kc_kc"] = (ms, 2.0 * s_ ** 3)
        del X, Y, Z
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cublas", action="store_true", help="add cuBLAS (torch FP64 matmul) timings of the same shapes")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--nb", type=int, default=1860)
    ap.add_argument("--naux", type=int, default=5500)
    ap.add_argument("--homo", type=int, default=179)
    ap.add_argument("--out", default="gpurun_out/contract_sweep.jsonl")
    ap.add_argument("--only", default="", help="comma-separated shape names")
    args = ap.parse_args()
    ctx = api.Context(0)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    cublas = cublas_reference(args.nb, args.naux, args.homo, args.reps) if args.cublas else {}
    with open(args.out, "w") as f:
        only = set(x for x in args.only.split(",") if x)
        for name, d, flops in shapes(args.nb, args.naux, args.homo):
            if only and name not in only:
                continue
            try:
                ms = api.contract_bench(ctx, d, args.reps)
                rec = {"shape": name, "M": d.M, "N": d.N, "K": d.K, "n_outer": d.n_outer, "n_batch": d.n_batch,
                       "lower": d.lower, "ms": round(ms, 4), "algorithmic_tflops": round(flops / ms * 1e-9, 3)}
                if name in cublas:
                    cms, cfl = cublas[name]
                    rec["cublas_ms"] = round(cms, 4)
                    rec["cublas_algorithmic_tflops"] = round(cfl / cms * 1e-9, 3)
                    rec["speedup_vs_cublas"] = round(cms / ms, 3)
                    if name == "epsilon_syrk":
                        rec["cublas_gemm_equivalent_tflops"] = round(cublas["epsilon_gemm_equivalent_tflops"], 3)
                        rec["cublas_what"] = "Ddgmm + Dgemm per occupied level (upstream OpenMP_CUDA::A_TDA), 8 levels timed"
            except Exception as e:  # noqa: BLE001
                rec = {"shape": name, "error": str(e)}
            print(json.dumps(rec), flush=True)
            f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
