// FP64 tensor-contraction engine for sm_100a: C = alpha * A * diag(d) * B^T-ish + beta * C
// on DMMA.8x8x4 (mma.sync.m8n8k4.f64), operands staged global->shared with a
// multi-stage cp.async pipeline into 128B-swizzled tiles (the same swizzle TMA's
// SWIZZLE_128B mode produces, so the consumer side is shared with the TMA path).
//
// Every GW-BSE contraction of SURVEY.md section 8a is an instance of this one kernel:
//   K1  M build           (T_P * C_m, C_n^T * W_P)        A K-contig, B K-contig
//   K2  aux rotation      (M[m] * R)                      A M-contig, B K-contig
//   K3  epsilon(w)        (A^T diag(d) A, two-level K)    A K-contig, B K-contig, d, lower-only
//   Sigma_x, BSE exchange / direct terms, Davidson projections.
//
// Index model.  A is addressed as  A.p + batch*A.s_batch + outer*A.s_outer + row*A.s_row + k*A.s_k
// with exactly one of (s_row, s_k) equal to 1; the contraction index is two-level:
// K_total = n_outer x K (each outer block is zero-padded to a multiple of 16).
// B likewise with row = output column.  C is written with arbitrary (c_sm, c_sn).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace xtpb {

struct GemmOperand {
  const double* p;
  long long s_row, s_k, s_outer, s_batch;
  // 1: the operand lies in a library allocation (which carries >= 256 bytes of slack behind its last element), so a
  // tile load may read up to 15 elements past the end of a k-row (single-box TMA loads of row-contiguous operands)
  int library_owned = 1;
};

struct GemmParams {
  GemmOperand A, B;
  double* C;
  long long c_sm, c_sn, c_batch;
  int c_n_inner;            // >0: two-level output column, col = co*c_n_inner + ci -> co*c_sn_outer + ci*c_sn
  long long c_sn_outer;
  int c_m_inner;            // >0: two-level output row,    row = ro*c_m_inner + ri -> ro*c_sm_outer + ri*c_sm
  long long c_sm_outer;
  const double* d;          // optional weights on the contraction index (nullptr = none)
  long long d_outer, d_batch;
  int M, N, K, n_outer, n_batch, splits;
  double alpha, beta;
  int lower;                // 1: skip tiles strictly above the diagonal (SYRK-style output)
  int a_vec, b_vec;         // 1: 16-byte cp.async legal for that operand
  int use_tma;              // 0: stay on the cp.async instance even when TMA is possible
  double* ws;               // split-K workspace: [batch][split][N][M] (M fastest)
};

constexpr int BK = 16;                 // doubles per k-tile: one 128-byte swizzle row
constexpr int ROW_BYTES = BK * 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void lds128(uint32_t addr, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(x), "=d"(y) : "r"(addr));
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---- shared-memory tile layouts (host+device so tests can check them on the CPU) ----
// K-contiguous operand: tile[row][16 k]; 16-byte chunk c of row r lives at chunk (c ^ (r & 7)).
__host__ __device__ constexpr inline uint32_t kc_offset(int row, int k) {
  return static_cast<uint32_t>(row * ROW_BYTES + ((((k >> 1) ^ (row & 7)) & 7) << 4) + ((k & 1) << 3));
}
// row-contiguous ("M-contiguous") operand: tile[row/16][k][16 rows]; chunk (r/2)%8 of k-row k at (chunk ^ (k & 7)).
__host__ __device__ constexpr inline uint32_t mc_offset(int row, int k) {
  return static_cast<uint32_t>((row >> 4) * (BK * ROW_BYTES) + k * ROW_BYTES +
                               (((((row >> 1) & 7) ^ (k & 7)) & 7) << 4) + ((row & 1) << 3));
}
// fragment slot -> tile row.  Slot i = lane >> 2.
//   K-contig : frag f covers rows 8f..8f+7,  slot i -> 8f + (i>>1) + 4(i&1)
//   M-contig : frags come in pairs covering 16 rows, slot i of frag f -> 16(f>>1) + 2i + (f&1)
__host__ __device__ constexpr inline int kc_slot_row(int f, int i) { return 8 * f + (i >> 1) + 4 * (i & 1); }
__host__ __device__ constexpr inline int mc_slot_row(int f, int i) { return 16 * (f >> 1) + 2 * i + (f & 1); }
// k consumed by k-slot t (= lane & 3) at half h (0/1), sub-step j (0/1) of a 16-wide k-tile
__host__ __device__ constexpr inline int slot_k(int h, int t, int j) { return 8 * h + 2 * t + j; }

template <int BM, int BN, int WM, int WN, bool A_KC, bool B_KC, int STAGES>
struct GemmCfg {
  static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
  static constexpr int THREADS = WARPS_M * WARPS_N * 32;
  static constexpr int MF = WM / 8, NF = WN / 8;
  static constexpr int A_BYTES = BM * ROW_BYTES, B_BYTES = BN * ROW_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES + ROW_BYTES;   // + d tile
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES;
  static_assert((BM * 8) % THREADS == 0 && (BN * 8) % THREADS == 0, "loader mapping");
  static_assert(MF % 2 == 0 && NF % 2 == 0, "fragment pairs");
};

// ---- per-thread loader with all index arithmetic hoisted out of the k-loop ----
// Chunk `c` (c < CHUNKS) of this thread:
//   K-contig : 16-byte column cc = tid & 7 (k = 2cc, 2cc+1) of row r0 + c*RSTEP;   r0 = tid >> 3, RSTEP = THREADS/8
//              (RSTEP is a multiple of 8, so the swizzle term (cc ^ (row & 7)) is the same for every chunk)
//   row-contig: row pair rg = tid % (R/2) (rows 2rg, 2rg+1) of k-row k0 + c*KSTEP; k0 = tid / (R/2), KSTEP = THREADS/(R/2)
// VEC: both operands allow 16-byte cp.async (compile-time, so the 8-byte path costs nothing when it is not needed).
template <int R, int THREADS, bool KC, bool VEC>
struct OperandLoader {
  static constexpr int CHUNKS = R * 8 / THREADS;
  static constexpr int RSTEP = THREADS / 8, KSTEP = THREADS / (R / 2);
  static_assert(RSTEP % 8 == 0 && THREADS % (R / 2) == 0, "chunk steps must preserve the swizzle phase");
  long long g_off0;           // element offset of chunk 0 from the tile base
  long long g_step;           // per-chunk increment (RSTEP rows / KSTEP k-rows)
  uint32_t s_off0;            // K-contig: smem offset of chunk 0;  row-contig: row-group part of the smem offset
  int first;                  // K-contig: k of the chunk (2cc);  row-contig: k-row of chunk 0
  int rows_ok;                // K-contig: rows_left - r0;  row-contig: valid bytes of the row pair (0, 8, 16)
  int rgsw;                   // row-contig: (rg & 7) swizzle phase
  __device__ __forceinline__ OperandLoader(long long s_row, long long s_k, int rows_left, int tid) {
    if (KC) {
      const int cc = tid & 7, r0 = tid >> 3;
      first = 2 * cc;
      rows_ok = rows_left - r0;                       // chunk c valid iff c*RSTEP < rows_ok
      g_off0 = (long long)r0 * s_row + first;
      g_step = (long long)RSTEP * s_row;
      s_off0 = (uint32_t)(r0 * ROW_BYTES + (((cc ^ (r0 & 7)) & 7) << 4));
      rgsw = 0;
    } else {
      const int rg = tid % (R / 2), k0 = tid / (R / 2), row = 2 * rg;
      first = k0;
      int nb = (rows_left - row) * 8;
      rows_ok = nb < 0 ? 0 : (nb > 16 ? 16 : nb);     // valid bytes of this row pair (constant over tiles)
      g_off0 = (long long)k0 * s_k + row;
      g_step = (long long)KSTEP * s_k;
      s_off0 = (uint32_t)((row >> 4) * (BK * ROW_BYTES));
      rgsw = rg & 7;
    }
  }
  // valid bytes along k of this thread's chunk column in a tile with k_left valid k values (K-contig only)
  __device__ __forceinline__ int k_bytes(int k_left) const {
    const int nb = (k_left - first) * 8;
    return nb < 0 ? 0 : (nb > 16 ? 16 : nb);
  }
  // issue chunk c of the tile whose first element is `base` into the stage operand area `sbase`;
  // kb = k_bytes(k_left) for K-contig operands, k_left itself for row-contig ones
  __device__ __forceinline__ void issue(int c, uint32_t sbase, const double* __restrict__ base, int kb) const {
    const double* src = base + g_off0 + c * g_step;
    if (KC) {
      const uint32_t dst = sbase + s_off0 + c * (RSTEP * ROW_BYTES);
      const int nb = c * RSTEP < rows_ok ? kb : 0;
      if (VEC) {
        cp_async16(dst, nb > 0 ? src : base, nb);
      } else {
        cp_async8(dst, nb >= 8 ? src : base, nb >= 8 ? 8 : 0);
        cp_async8(dst + 8, nb >= 16 ? src + 1 : base, nb >= 16 ? 8 : 0);
      }
    } else {
      const int k = first + c * KSTEP;
      const uint32_t dst = sbase + s_off0 + k * ROW_BYTES + (((rgsw ^ (k & 7)) & 7) << 4);
      const int nb = k < kb ? rows_ok : 0;
      if (VEC) {
        cp_async16(dst, nb > 0 ? src : base, nb);
      } else {
        cp_async8(dst, nb >= 8 ? src : base, nb >= 8 ? 8 : 0);
        cp_async8(dst + 8, nb >= 16 ? src + 1 : base, nb >= 16 ? 8 : 0);
      }
    }
  }
};

// ---- fragment addressing, computed once per thread ---------------------------------------------------------
// K-contig operand: fragment f, half h:  (w0 + rl)*128 + (((4h + lt) ^ rl) << 4) + f*1024, rl = row & 7 of the slot
// row-contig operand: pair q, half h, sub-step j: (w0/16 + q)*2048 + (8h + 2lt + j)*128 + ((li ^ (2lt + j)) << 4)
// -> two per-thread bases per operand (index h for K-contig, j for row-contig) plus compile-time immediates.
template <bool KC>
__device__ __forceinline__ void frag_offsets(uint32_t (&off)[2], int w0, int li, int lt) {
  const int rl = (li >> 1) + 4 * (li & 1);
#pragma unroll
  for (int x = 0; x < 2; ++x)
    off[x] = KC ? (uint32_t)((w0 + rl) * ROW_BYTES + ((((4 * x + lt) ^ rl) & 7) << 4))
                : (uint32_t)((w0 >> 4) * (BK * ROW_BYTES) + (2 * lt + x) * ROW_BYTES + (((li ^ (2 * lt + x)) & 7) << 4));
}

// ---- one 16-wide k-tile of DMMAs for this warp -------------------------------------------------------------
// The 2*MF "steps" (half h = 8 k values, one A fragment load feeding 2*NF DMMAs each) are software-pipelined by
// hand: A fragments are double-buffered in registers and loaded one step ahead, the B fragments (and their chi0
// weights) of the second half are fetched and scaled while the first half is still issuing DMMAs.  `hook(s)` is
// called once per step so the caller can spread its producer work (cp.async chunks or the TMA issue) between the
// DMMA groups; no warp ever sits in a long non-tensor instruction run.
template <typename Cfg, bool A_KC, bool B_KC, bool HAS_D, typename Hook>
__device__ __forceinline__ void consume_tile(double (&acc)[Cfg::MF][Cfg::NF][2], uint32_t sA, uint32_t sB, uint32_t sD,
                                             const uint32_t (&a_off)[2], const uint32_t (&b_off)[2], uint32_t d_off,
                                             Hook&& hook) {
  constexpr int STEPS = 2 * Cfg::MF;
  double bf[2][Cfg::NF][2];
  double af[2][2];
  auto load_b = [&](int h, double (&b)[Cfg::NF][2]) {
    if (B_KC) {
#pragma unroll
      for (int nf = 0; nf < Cfg::NF; ++nf) lds128(sB + b_off[h] + nf * (8 * ROW_BYTES), b[nf][0], b[nf][1]);
    } else {
#pragma unroll
      for (int np = 0; np < Cfg::NF / 2; ++np)
#pragma unroll
        for (int j = 0; j < 2; ++j)
          lds128(sB + b_off[j] + h * (8 * ROW_BYTES) + np * (BK * ROW_BYTES), b[2 * np][j], b[2 * np + 1][j]);
    }
  };
  auto scale_b = [&](double d0, double d1, double (&b)[Cfg::NF][2]) {
#pragma unroll
    for (int nf = 0; nf < Cfg::NF; ++nf) {
      b[nf][0] *= d0;
      b[nf][1] *= d1;
    }
  };
  // A fragment of step s = h*MF + i:  K-contig: i = mf, (a0,a1) = two k sub-steps of fragment mf;
  //                                   M-contig: i = 2*mp + j, (a0,a1) = fragments 2mp, 2mp+1 at k sub-step j
  auto load_a = [&](int s, double (&a)[2]) {
    const int h = s / Cfg::MF, i = s % Cfg::MF;
    if (A_KC) lds128(sA + a_off[h] + i * (8 * ROW_BYTES), a[0], a[1]);
    else lds128(sA + a_off[i & 1] + h * (8 * ROW_BYTES) + (i >> 1) * (BK * ROW_BYTES), a[0], a[1]);
  };

  double dn0 = 1.0, dn1 = 1.0;
  load_b(0, bf[0]);
  if (HAS_D) {
    double d0, d1;
    lds128(sD + d_off, d0, d1);
    scale_b(d0, d1, bf[0]);
  }
  load_a(0, af[0]);
#pragma unroll
  for (int s = 0; s < STEPS; ++s) {
    const int h = s / Cfg::MF, i = s % Cfg::MF;
    if (s + 1 < STEPS) load_a(s + 1, af[(s + 1) & 1]);   // one step (2*NF DMMAs) ahead of its use
    if (s == 0) {
      load_b(1, bf[1]);
      if (HAS_D) lds128(sD + d_off + 64, dn0, dn1);
    }
    hook(s);
    const double a0 = af[s & 1][0], a1 = af[s & 1][1];
    if (A_KC) {
#pragma unroll
      for (int nf = 0; nf < Cfg::NF; ++nf) dmma884(acc[i][nf][0], acc[i][nf][1], a0, bf[h][nf][0]);
#pragma unroll
      for (int nf = 0; nf < Cfg::NF; ++nf) dmma884(acc[i][nf][0], acc[i][nf][1], a1, bf[h][nf][1]);
    } else {
      const int mp = i >> 1, j = i & 1;
#pragma unroll
      for (int nf = 0; nf < Cfg::NF; ++nf) dmma884(acc[2 * mp][nf][0], acc[2 * mp][nf][1], a0, bf[h][nf][j]);
#pragma unroll
      for (int nf = 0; nf < Cfg::NF; ++nf)
        dmma884(acc[2 * mp + 1][nf][0], acc[2 * mp + 1][nf][1], a1, bf[h][nf][j]);
    }
    if (HAS_D && s == Cfg::MF / 2) scale_b(dn0, dn1, bf[1]);
  }
}

// ---- epilogue: registers -> C (or the split-K workspace) through the two-level output maps ----
template <typename Cfg, bool A_KC, bool B_KC>
__device__ __forceinline__ void store_tile(const GemmParams& p, const double (&acc)[Cfg::MF][Cfg::NF][2], int m0, int n0,
                                           int wm0, int wn0, int li, int lt, int split, int batch) {
  const bool to_ws = p.splits > 1;
  double* Cb = to_ws ? p.ws + ((long long)batch * p.splits + split) * (long long)p.M * p.N
                     : p.C + (long long)batch * p.c_batch;
  const long long csm = to_ws ? 1 : p.c_sm, csn = to_ws ? (long long)p.M : p.c_sn;
  const double alpha = to_ws ? 1.0 : p.alpha, beta = to_ws ? 0.0 : p.beta;
#pragma unroll
  for (int mf = 0; mf < Cfg::MF; ++mf) {
    const int r = m0 + wm0 + (A_KC ? kc_slot_row(mf, li) : mc_slot_row(mf, li));
    if (r >= p.M) continue;
    const long long roff = (!to_ws && p.c_m_inner > 0)
                               ? (long long)(r / p.c_m_inner) * p.c_sm_outer + (long long)(r % p.c_m_inner) * csm
                               : (long long)r * csm;
#pragma unroll
    for (int nf = 0; nf < Cfg::NF; ++nf)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int s = 2 * lt + e;
        const int c = n0 + wn0 + (B_KC ? kc_slot_row(nf, s) : mc_slot_row(nf, s));
        if (c >= p.N) continue;
        if (p.lower && c > r) continue;
        const long long coff = (!to_ws && p.c_n_inner > 0)
                                   ? (long long)(c / p.c_n_inner) * p.c_sn_outer + (long long)(c % p.c_n_inner) * csn
                                   : (long long)c * csn;
        double* dst = Cb + roff + coff;
        double v = alpha * acc[mf][nf][e];
        if (beta != 0.0) v += beta * (*dst);
        *dst = v;
      }
  }
}

// ============================================================ cp.async instance (any alignment, any strides)
// Output tile of a CTA.  CTAs are launched in blockIdx order, so the ~148 that run together should share operand
// panels: tiles are walked in groups of kTileGroup tile columns, row by row inside a group (a wave then touches about
// 8 + 18 panels of a large product instead of all tile rows + 3 columns; ncu of the eps SYRK at C60 size showed 13x the
// compulsory DRAM reads with the plain column-major walk, profiles/r02_contract_tma_ncu_c60.md).
constexpr int kTileGroup = 8;
__device__ __forceinline__ void tile_of_block(int block, int tiles_m, int tiles_n, int& tm, int& tn) {
  const int per_group = kTileGroup * tiles_m;
  const int grp = block / per_group;
  const int n_first = grp * kTileGroup;
  const int gsz = min(tiles_n - n_first, kTileGroup);
  const int r = block - grp * per_group;
  tn = n_first + r % gsz;
  tm = r / gsz;
}

template <typename Cfg, int BM, int BN, int WM, int WN, bool A_KC, bool B_KC, int STAGES, bool HAS_D, bool VEC>
__global__ void __launch_bounds__(Cfg::THREADS) contract_kernel(const GemmParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t smem = smem_u32(smem_raw);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp % Cfg::WARPS_M) * WM, wn0 = (warp / Cfg::WARPS_M) * WN;
  const int li = lane >> 2, lt = lane & 3;

  const int tiles_m = (p.M + BM - 1) / BM;
  int tm, tn;
  tile_of_block(blockIdx.x, tiles_m, gridDim.x / tiles_m, tm, tn);
  if (p.lower && tn * BN > tm * BM + BM - 1) return;
  const int split = blockIdx.z % p.splits, batch = blockIdx.z / p.splits;

  const int tpo = (p.K + BK - 1) / BK;                 // k-tiles per outer block
  const int nkt_total = tpo * p.n_outer;
  const int per = (nkt_total + p.splits - 1) / p.splits;
  const int kt_begin = split * per;
  const int kt_end = min(nkt_total, kt_begin + per);
  const int nkt = max(0, kt_end - kt_begin);

  const int m0 = tm * BM, n0 = tn * BN;
  const double* Ab = p.A.p + (long long)batch * p.A.s_batch + (long long)m0 * p.A.s_row;
  const double* Bb = p.B.p + (long long)batch * p.B.s_batch + (long long)n0 * p.B.s_row;
  const double* Db = HAS_D ? p.d + (long long)batch * p.d_batch : nullptr;
  const int m_left = p.M - m0, n_left = p.N - n0;

  // loader state, computed once per thread: only the tile base pointer changes from tile to tile
  const OperandLoader<BM, Cfg::THREADS, A_KC, VEC> ldA(p.A.s_row, p.A.s_k, m_left, tid);
  const OperandLoader<BN, Cfg::THREADS, B_KC, VEC> ldB(p.B.s_row, p.B.s_k, n_left, tid);
  constexpr int NA = OperandLoader<BM, Cfg::THREADS, A_KC, VEC>::CHUNKS;
  constexpr int NB = OperandLoader<BN, Cfg::THREADS, B_KC, VEC>::CHUNKS;

  // tile -> (outer, k0) walked incrementally (no division in the loop)
  int l_outer = kt_begin / tpo, l_k0 = (kt_begin - l_outer * tpo) * BK;
  auto advance_tile = [&]() {
    l_k0 += BK;
    if (l_k0 >= p.K) { l_k0 = 0; ++l_outer; }
  };
  auto load_tile_full = [&](int stage) {      // prologue: whole tile at once
    const uint32_t sA = smem + stage * Cfg::STAGE_BYTES, sB = sA + Cfg::A_BYTES, sD = sB + Cfg::B_BYTES;
    const int k_left = p.K - l_k0;
    const double* gA = Ab + (long long)l_outer * p.A.s_outer + (long long)l_k0 * p.A.s_k;
    const double* gB = Bb + (long long)l_outer * p.B.s_outer + (long long)l_k0 * p.B.s_k;
    const int kbA = A_KC ? ldA.k_bytes(k_left) : k_left, kbB = B_KC ? ldB.k_bytes(k_left) : k_left;
#pragma unroll
    for (int c = 0; c < NA; ++c) ldA.issue(c, sA, gA, kbA);
#pragma unroll
    for (int c = 0; c < NB; ++c) ldB.issue(c, sB, gB, kbB);
    if (HAS_D && tid < BK) {
      const bool v = tid < k_left;
      const double* src = v ? Db + (long long)l_outer * p.d_outer + l_k0 + tid : Db;
      cp_async8(sD + tid * 8, src, v ? 8 : 0);
    }
  };

  double acc[Cfg::MF][Cfg::NF][2];
#pragma unroll
  for (int i = 0; i < Cfg::MF; ++i)
#pragma unroll
    for (int j = 0; j < Cfg::NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nkt) {
      load_tile_full(s);
      advance_tile();
    }
    cp_async_commit();
  }

  uint32_t a_off[2], b_off[2];
  frag_offsets<A_KC>(a_off, wm0, li, lt);
  frag_offsets<B_KC>(b_off, wn0, li, lt);
  const uint32_t d_off = (uint32_t)(2 * lt * 8);

  // Main loop: one barrier per k-tile; the cp.async traffic of the tile STAGES-1 ahead is spread over the steps.
  constexpr int NCH = NA + NB, STEPS = 2 * Cfg::MF;
#pragma unroll 1
  for (int it = 0; it < nkt; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const int nxt = it + STAGES - 1;
    const bool do_load = nxt < nkt;
    const uint32_t lA = smem + (nxt % STAGES) * Cfg::STAGE_BYTES, lB = lA + Cfg::A_BYTES, lD = lB + Cfg::B_BYTES;
    const int l_kleft = p.K - l_k0;
    const double* gA = Ab + (long long)l_outer * p.A.s_outer + (long long)l_k0 * p.A.s_k;
    const double* gB = Bb + (long long)l_outer * p.B.s_outer + (long long)l_k0 * p.B.s_k;
    const double* gD = HAS_D ? Db + (long long)l_outer * p.d_outer + l_k0 : nullptr;
    const int kbA = A_KC ? ldA.k_bytes(l_kleft) : l_kleft, kbB = B_KC ? ldB.k_bytes(l_kleft) : l_kleft;
    const int stage = it % STAGES;
    const uint32_t sA = smem + stage * Cfg::STAGE_BYTES, sB = sA + Cfg::A_BYTES, sD = sB + Cfg::B_BYTES;
    consume_tile<Cfg, A_KC, B_KC, HAS_D>(acc, sA, sB, sD, a_off, b_off, d_off, [&](int s) {
      if (do_load) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          if (c * STEPS / NCH != s) continue;
          if (c < NA) ldA.issue(c, lA, gA, kbA);
          else ldB.issue(c - NA, lB, gB, kbB);
        }
        if (HAS_D && s == STEPS - 1 && tid < BK) {
          const bool v = tid < l_kleft;
          cp_async8(lD + tid * 8, v ? gD + tid : Db, v ? 8 : 0);
        }
      }
    });
    cp_async_commit();
    if (do_load) advance_tile();
  }
  cp_async_wait<0>();
  store_tile<Cfg, A_KC, B_KC>(p, acc, m0, n0, wm0, wn0, li, lt, split, batch);
}

// ============================================================ TMA instance (16-byte aligned operands)
// Operand tiles arrive through cp.async.bulk.tensor (SWIZZLE_128B produces exactly the kc/mc tile layouts above;
// out-of-range rows and the k tail of every outer block are zero-filled by the hardware), completion is tracked by
// one "full" mbarrier per stage, and consumers hand a stage back through an "empty" mbarrier: there is no
// __syncthreads and no address arithmetic in the main loop, and the warps of a CTA may drift up to STAGES-1 tiles
// apart, which decouples their fragment-load and DMMA phases.  Lane 0 of warp 0 issues the copies for the tile
// STAGES-1 ahead in the middle of its own DMMA stream (a dozen instructions per tile).
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.b32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 24)) __trap();     // a lost arrival becomes an error, never a hang
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2, int c3,
                                            int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

struct TmaCoords {        // per-operand multipliers: 0 when the tensor map has no outer / batch dimension
  int a_outer, a_batch, b_outer, b_batch;
  int a_mc5, b_mc5;       // row-contiguous operand described by the 5-D map (one box per tile)
};

template <typename Cfg, int BM, int BN, int WM, int WN, bool A_KC, bool B_KC, int STAGES, bool HAS_D>
__global__ void __launch_bounds__(Cfg::THREADS) contract_tma_kernel(const GemmParams p, const TmaCoords tc,
                                                                    const __grid_constant__ CUtensorMap mapA,
                                                                    const __grid_constant__ CUtensorMap mapB) {
  // stage = [A tile | B tile | d tile padded to 1 KiB]: every operand tile starts on a 1 KiB boundary (swizzle atom)
  constexpr int TSTAGE = Cfg::A_BYTES + Cfg::B_BYTES + 1024;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t smem = smem_u32(smem_raw);
  const uint32_t bar_full = smem + STAGES * TSTAGE, bar_empty = bar_full + STAGES * 8;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp % Cfg::WARPS_M) * WM, wn0 = (warp / Cfg::WARPS_M) * WN;
  const int li = lane >> 2, lt = lane & 3;

  const int tiles_m = (p.M + BM - 1) / BM;
  int tm, tn;
  tile_of_block(blockIdx.x, tiles_m, gridDim.x / tiles_m, tm, tn);
  if (p.lower && tn * BN > tm * BM + BM - 1) return;
  const int split = blockIdx.z % p.splits, batch = blockIdx.z / p.splits;

  const int tpo = (p.K + BK - 1) / BK;
  const int nkt_total = tpo * p.n_outer;
  const int per = (nkt_total + p.splits - 1) / p.splits;
  const int kt_begin = split * per;
  const int kt_end = min(nkt_total, kt_begin + per);
  const int nkt = max(0, kt_end - kt_begin);
  const int m0 = tm * BM, n0 = tn * BN;
  const double* Db = HAS_D ? p.d + (long long)batch * p.d_batch : nullptr;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, HAS_D ? 1 + BK : 1);       // expect_tx arrival (+ one per d element copy)
      mbar_init(bar_empty + 8 * s, Cfg::THREADS / 32);       // one arrival per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  int l_outer = kt_begin / tpo, l_k0 = (kt_begin - l_outer * tpo) * BK;
  auto advance_tile = [&]() {
    l_k0 += BK;
    if (l_k0 >= p.K) { l_k0 = 0; ++l_outer; }
  };
  // warp 0 only: copies of tile (l_outer, l_k0) into `stage`
  auto produce = [&](int stage) {
    const uint32_t sA = smem + stage * TSTAGE, sB = sA + Cfg::A_BYTES, sD = sB + Cfg::B_BYTES;
    const uint32_t full = bar_full + 8 * stage;
    if (HAS_D && lane < BK) {
      const bool v = l_k0 + lane < p.K;
      cp_async8(sD + lane * 8, v ? Db + (long long)l_outer * p.d_outer + l_k0 + lane : Db, v ? 8 : 0);
      mbar_cp_async_arrive_noinc(full);
    }
    if (lane == 0) {
      mbar_arrive_expect_tx(full, Cfg::A_BYTES + Cfg::B_BYTES);
      if (A_KC) {
        tma_load_4d(sA, &mapA, full, l_k0, m0, l_outer * tc.a_outer, batch * tc.a_batch);
      } else if (tc.a_mc5) {
        tma_load_5d(sA, &mapA, full, 0, l_k0, m0 >> 4, l_outer * tc.a_outer, batch * tc.a_batch);
      } else {
#pragma unroll
        for (int g = 0; g < BM / 16; ++g)
          tma_load_4d(sA + g * (BK * ROW_BYTES), &mapA, full, m0 + 16 * g, l_k0, l_outer * tc.a_outer, batch * tc.a_batch);
      }
      if (B_KC) {
        tma_load_4d(sB, &mapB, full, l_k0, n0, l_outer * tc.b_outer, batch * tc.b_batch);
      } else if (tc.b_mc5) {
        tma_load_5d(sB, &mapB, full, 0, l_k0, n0 >> 4, l_outer * tc.b_outer, batch * tc.b_batch);
      } else {
#pragma unroll
        for (int g = 0; g < BN / 16; ++g)
          tma_load_4d(sB + g * (BK * ROW_BYTES), &mapB, full, n0 + 16 * g, l_k0, l_outer * tc.b_outer, batch * tc.b_batch);
      }
    }
  };

  double acc[Cfg::MF][Cfg::NF][2];
#pragma unroll
  for (int i = 0; i < Cfg::MF; ++i)
#pragma unroll
    for (int j = 0; j < Cfg::NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll 1
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nkt) {
      if (warp == 0) produce(s);
      advance_tile();
    }
  }

  uint32_t a_off[2], b_off[2];
  frag_offsets<A_KC>(a_off, wm0, li, lt);
  frag_offsets<B_KC>(b_off, wn0, li, lt);
  const uint32_t d_off = (uint32_t)(2 * lt * 8);
  constexpr int STEPS = 2 * Cfg::MF;

#pragma unroll 1
  for (int it = 0; it < nkt; ++it) {
    const int stage = it % STAGES;
    const int nxt = it + STAGES - 1;
    const bool do_load = nxt < nkt;
    mbar_wait(bar_full + 8 * stage, (uint32_t)((it / STAGES) & 1));
    const uint32_t sA = smem + stage * TSTAGE, sB = sA + Cfg::A_BYTES, sD = sB + Cfg::B_BYTES;
    consume_tile<Cfg, A_KC, B_KC, HAS_D>(acc, sA, sB, sD, a_off, b_off, d_off, [&](int s) {
      if (s == STEPS / 2 && warp == 0 && do_load) {
        const int ns = nxt % STAGES;       // consumed as tile it-1: wait until every warp has handed it back
        if (it > 0 && lane == 0) mbar_wait(bar_empty + 8 * ns, (uint32_t)(((it - 1) / STAGES) & 1));
        __syncwarp();
        produce(ns);
      }
    });
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8 * stage);
    if (do_load) advance_tile();
  }
  store_tile<Cfg, A_KC, B_KC>(p, acc, m0, n0, wm0, wn0, li, lt, split, batch);
}

}  // namespace xtpb
