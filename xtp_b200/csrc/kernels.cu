// Element-wise and reduction kernels of the GW-BSE path (everything that is not a contraction):
// chi0 weights, Sigma_c plasmon-pole sums (FP64 ALU / HBM bound), BSE diagonal, Davidson vector ops.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "internal.h"

namespace xtpb {

namespace {

constexpr double kFourPi = 12.566370614359172953850573533118;

inline int blocks_for(long long n, int threads, int cap = 65535 * 16) {
  long long b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

#define LAUNCH_CHECK()                 \
  do {                                 \
    XTPB_CUDA(cudaGetLastError());     \
    ++g_launch_count;                  \
  } while (0)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum, result valid in thread 0 (deterministic order)
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x < 32) {
    r = threadIdx.x < THREADS / 32 ? sh[threadIdx.x] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

// FP64 reciprocal without the library slow path: MUFU.RCP64H seed (rel. error <= 2^-23, PTX rcp.approx.ftz.f64)
// refined by one cubic step r(1 + e + e^2), e = 1 - x r  ->  rel. error ~2^-69 + rounding (3 DFMA, <= 2 ulp).
// Valid for finite normal x; x = 0 gives inf (callers route |x| < 0.25 through the damped branch).
__device__ __forceinline__ double rcp_fast(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);
  const double t = fma(e, e, e);
  return fma(r, t, r);
}

// sin(2 pi x) for |x| <= 0.25 (argument in [-pi/2, pi/2]): odd Taylor polynomial through theta^21
// (truncation < 2e-18), Horner in theta^2 -- 12 FP64 operations, no range reduction, no slow path.
__device__ __forceinline__ double sin2pi_quarter(double x) {
  const double t = 6.283185307179586476925286766559 * x, t2 = t * t;
  double p = -1.9572941063391261230e-20;        // -1/21!
  p = fma(p, t2, 8.2206352466243297170e-18);    //  1/19!
  p = fma(p, t2, -2.8114572543455207632e-15);   // -1/17!
  p = fma(p, t2, 7.6471637318198164759e-13);    //  1/15!
  p = fma(p, t2, -1.6059043836821614599e-10);   // -1/13!
  p = fma(p, t2, 2.5052108385441718775e-08);    //  1/11!
  p = fma(p, t2, -2.7557319223985890653e-06);   // -1/9!
  p = fma(p, t2, 1.9841269841269841270e-04);    //  1/7!
  p = fma(p, t2, -8.3333333333333333333e-03);   // -1/5!
  p = fma(p, t2, 1.6666666666666666667e-01);    //  1/3!  (sign folded below)
  return fma(-t * t2, p, t);                    // t - t^3 (1/3! - t^2/5! + ...)
}

// damped branch of the Rohlfing-stabilised inverse: 0.5 (1 - cos 4 pi x) / x = sin^2(2 pi x) / x  (no cancellation);
// r = 1/x is already at hand; Taylor limit 4 pi^2 x below the range where r is finite.
__device__ __forceinline__ double ppm_ginv_damped(double x, double r) {
  const double s = sin2pi_quarter(x);
  return fabs(x) < 1e-100 ? (0.25 * kFourPi * kFourPi) * x : s * s * r;
}

// The damped branch without a reciprocal: sin^2(2 pi x)/x = x h(x^2), h(u) = sum_{k>=1} (-1)^(k+1) (4 pi)^(2k) u^(k-1) / (2 (2k)!)
// (entire function; 13 terms leave < 1e-15 relative for u <= 1/16, checked against 40-digit arithmetic) -- 14 FP64
// operations, x = 0 gives 0.  Used where whole warps sit inside the damping window (compressed grid scan).
__device__ __forceinline__ double ppm_ginv_poly(double x) {
  const double u = x * x;
  double h = 4.7076855031461505630e+01;
  h = fma(h, u, -1.9377648362907347975e+02);
  h = fma(h, u, 6.7736136257549954211e+02);
  h = fma(h, u, -1.9817217134061364861e+03);
  h = fma(h, u, 4.7687717542357461594e+03);
  h = fma(h, u, -9.2407715743593665354e+03);
  h = fma(h, u, 1.4044288705238401560e+04);
  h = fma(h, u, -1.6186442488460224111e+04);
  h = fma(h, u, 1.3530243473087691496e+04);
  h = fma(h, u, -7.7113140956002125264e+03);
  h = fma(h, u, 2.7346181506141992876e+03);
  h = fma(h, u, -5.1951515218134633193e+02);
  h = fma(h, u, 3.9478417604357434475e+01);
  return x * h;
}

// |x| < 0.25, read from the exponent word on the integer pipe (the FP64 pipe only carries subtract/refine/accumulate)
__device__ __forceinline__ bool ppm_in_window(double x) {
  return (static_cast<unsigned>(__double2hiint(x)) & 0x7fffffffu) < 0x3fd00000u;
}

// Rohlfing-stabilised inverse, upstream Sigma_PPM::Stabilize (sigma_ppm.cc): 1/x for |x| >= 0.25,
// 0.5 (1 - cos 4 pi x) / x otherwise (-> 0 at x = 0).
__device__ __forceinline__ double ppm_ginv(double x) {
  const double r = rcp_fast(x);
  return ppm_in_window(x) ? ppm_ginv_damped(x, r) : r;
}

// ------------------------------------------------------------------ chi0 weights
// d[w][m][k], k <-> level a = a0 + k (relative to rpamin); zero for a < n_occ (alignment padding).
// Upstream: the `denom` vector of RPA::calculate_epsilon<imag> (rpa.cc).
// e_m: energies by first index (all levels); e_n: energies of the second-index columns held by this rank
// (n_occ_n of them occupied); identical arrays on a single GPU.
__global__ void chi0_weights_kernel(double* __restrict__ d, const double* __restrict__ e_m,
                                    const double* __restrict__ e_n, int n_occ, int n_occ_n, int a0, int K,
                                    const double* __restrict__ omegas, int n_omega, int imag, double eta) {
  const long long total = (long long)n_omega * n_occ * K;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % K);
    const int m = (int)((idx / K) % n_occ);
    const int w = (int)(idx / ((long long)K * n_occ));
    const int a = a0 + k;
    double v = 0.0;
    if (a >= n_occ_n) {
      const double dE = e_n[a] - e_m[m];
      const double om = omegas[w];
      if (imag) {
        v = 4.0 * dE / (dE * dE + om * om);
      } else {
        const double dm = dE - om, dp = dE + om, eta2 = eta * eta;
        v = 2.0 * (dm / (dm * dm + eta2) + dp / (dp * dp + eta2));
      }
    }
    d[idx] = v;
  }
}

__global__ void set_identity_kernel(double* A, int n, long long ld) {
  const long long total = (long long)n * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % n), c = (int)(idx / n);
    A[r + c * ld] = r == c ? 1.0 : 0.0;
  }
}
__global__ void add_diagonal_kernel(double* A, int n, long long ld, double v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) A[i + i * ld] += v;
}
__global__ void scale_columns_kernel(double* A, int rows, int cols, long long ld, const double* __restrict__ scale) {
  const long long total = (long long)rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % rows), c = (int)(idx / rows);
    A[r + c * ld] *= scale[c];
  }
}
__global__ void zero_strict_lower_kernel(double* A, int n, long long ld) {
  const long long total = (long long)n * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % n), j = (int)(idx / n);
    if (i > j) A[i + (long long)j * ld] = 0.0;
  }
}
__global__ void extract_diagonal_kernel(const double* A, int n, long long ld, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = A[i + i * ld];
}
__global__ void copy_2d_kernel(double* dst, long long ldd, const double* __restrict__ src, long long lds, int rows,
                               long long cols) {
  const long long total = (long long)rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % rows);
    const long long c = idx / rows;
    dst[r + c * ldd] = src[r + c * lds];
  }
}
__global__ void scale_kernel(double* x, long long n, double a) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] *= a;
}
__global__ void axpby_kernel(double* y, const double* __restrict__ x, long long n, double a, double b) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = a * x[i] + (b == 0.0 ? 0.0 : b * y[i]);
}
__global__ void extract_window_kernel(double* __restrict__ dst, long long dst_ld, long long dst_slab,
                                      const double* __restrict__ M, long long ldn, long long slab, int m0, int mcnt,
                                      int n0, int ncnt, int naux, const double* __restrict__ scale) {
  const long long total = (long long)mcnt * naux * ncnt;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % ncnt);
    const int P = (int)((idx / ncnt) % naux);
    const int i = (int)(idx / ((long long)ncnt * naux));
    double v = M[(long long)(m0 + i) * slab + (long long)P * ldn + n0 + j];
    if (scale) v *= scale[P];
    dst[(long long)i * dst_slab + (long long)P * dst_ld + j] = v;
  }
}

// ------------------------------------------------------------------ packed symmetric AO slices -> full
// One 32x32 tile (ti >= tj) of the lower triangle per block: rows of the packed triangle are read coalesced, written
// to the lower tile directly and to the mirrored upper tile through a shared-memory transpose.  HBM-bound:
// 4 n^2 bytes read + 8 n^2 written per slice.
__global__ void __launch_bounds__(256) unpack_symmetric_kernel(double* __restrict__ full, long long ld,
                                                               long long full_slice, const double* __restrict__ packed,
                                                               long long pk_slice, int n) {
  __shared__ double tile[32][33];
  // linear block index -> (ti, tj) with ti >= tj
  const int t = blockIdx.x;
  int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((long long)(ti + 1) * (ti + 2) / 2 <= t) ++ti;
  while ((long long)ti * (ti + 1) / 2 > t) --ti;
  const int tj = t - ti * (ti + 1) / 2;
  const double* src = packed + (long long)blockIdx.y * pk_slice;
  double* dst = full + (long long)blockIdx.y * full_slice;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int mu = ti * 32 + r, nu = tj * 32 + tx;
    double v = 0.0;
    if (mu < n && nu <= mu) v = src[(long long)mu * (mu + 1) / 2 + nu];
    tile[r][tx] = v;
    // column-major full matrix: element (row mu, col nu) at mu + nu*ld; writing row-index-fastest needs the transpose,
    // so the direct write below covers (row nu, col mu) = upper mirror, contiguous in nu
    if (mu < n && nu <= mu) dst[nu + (long long)mu * ld] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    // transposed read: element (mu = ti*32 + tx, nu = tj*32 + r) -> (row mu, col nu), contiguous in mu
    const int mu = ti * 32 + tx, nu = tj * 32 + r;
    if (mu < n && nu < mu) dst[mu + (long long)nu * ld] = tile[tx][r];
  }
}

// ------------------------------------------------------------------ Sigma_c, plasmon-pole model
// Upstream Sigma_PPM::CalcCorrelationDiagElement (sigma_ppm.cc):
//   Sigma_c(level, w) = sum_P fac_P sum_m M~[level](m,P)^2 * ginv(w - e_m +/- Omega_P)   (+ occupied, - unoccupied)
//
// (1) grid kernel: the QP grid solver (GW::SolveQP_Grid, gw.cc) needs ~1001 frequencies per level.  Threads own
//     frequencies (NW each, held in registers, consecutive lanes = consecutive grid points so the |x|<0.25 branch
//     is warp-coherent), the (P,m) elements of the level's slab stream through shared memory and are broadcast.
//     Bound: FP64 ALU (one reciprocal per element and frequency); the slab is read once per 2*T*NW frequencies.
constexpr int kGridThreads = 128, kGridNW = 4, kGridTile = 512, kGridEl = 2;

// Per pole and frequency: 1 subtract, 1 MUFU.RCP64H seed + 3 DFMA refinement, 1 DFMA accumulate, and an integer
// window test; the reciprocals of the kGridEl*NW evaluations of one step are independent chains.  If any of them
// falls inside the damping window the (rare, warp-coherent: lanes are consecutive grid points) branch replaces those
// reciprocals by sin^2(2 pi x)/x before the accumulate -- no selects on the common path.
// blockIdx.z splits the aux range so that small level counts still fill the machine; partial sums are reduced in a
// fixed order by sigma_ppm_grid_reduce (deterministic).
__global__ void __launch_bounds__(kGridThreads) sigma_ppm_grid_kernel(
    const double* __restrict__ M, long long ldn, long long slab, int ntotal, int naux, int n_occ,
    const double* __restrict__ energies, const double* __restrict__ ppm_freq, const double* __restrict__ ppm_fac,
    const int* __restrict__ level_slab, const double* __restrict__ omega0, double domega, int n_omega,
    double* __restrict__ out, long long out_split_stride) {
  __shared__ double2 tile[kGridTile];
  static_assert(kGridTile % kGridEl == 0, "tile is consumed kGridEl poles at a time");
  const int level = blockIdx.y;
  const double* S = M + (long long)level_slab[level] * slab;
  const int jbase = blockIdx.x * (kGridThreads * kGridNW) + threadIdx.x;
  const int p_per = (naux + gridDim.z - 1) / gridDim.z;
  const int p_begin = blockIdx.z * p_per, p_end = min(naux, p_begin + p_per);
  double om[kGridNW], acc[kGridNW];
#pragma unroll
  for (int w = 0; w < kGridNW; ++w) {
    om[w] = omega0[level] + domega * (double)(jbase + w * kGridThreads);
    acc[w] = 0.0;
  }
  for (int P = p_begin; P < p_end; ++P) {
    const double fac = ppm_fac[P];
    if (fac == 0.0) continue;     // uniform across the block
    const double Om = ppm_freq[P];
    const double* row = S + (long long)P * ldn;
    for (int m0 = 0; m0 < ntotal; m0 += kGridTile) {
      __syncthreads();
      for (int t = threadIdx.x; t < kGridTile; t += kGridThreads) {
        const int m = m0 + t;
        double2 el = make_double2(0.0, -1.0e30);            // padding: weight 0, far away
        if (m < ntotal) {
          const double v = row[m];
          el.x = fac * v * v;
          el.y = energies[m] + (m < n_occ ? -Om : Om);     // pole position z: x = w - z
        }
        tile[t] = el;
      }
      __syncthreads();
      const int cnt = (min(kGridTile, ntotal - m0) + kGridEl - 1) / kGridEl * kGridEl;
#pragma unroll 2
      for (int t = 0; t < cnt; t += kGridEl) {
        double2 el[kGridEl];
        double x[kGridEl][kGridNW], r[kGridEl][kGridNW];
        bool any = false;
#pragma unroll
        for (int g = 0; g < kGridEl; ++g) {
          el[g] = tile[t + g];
#pragma unroll
          for (int w = 0; w < kGridNW; ++w) {
            x[g][w] = om[w] - el[g].y;
            r[g][w] = rcp_fast(x[g][w]);
            any |= ppm_in_window(x[g][w]);
          }
        }
        if (any) {
#pragma unroll
          for (int g = 0; g < kGridEl; ++g)
#pragma unroll
            for (int w = 0; w < kGridNW; ++w)
              if (ppm_in_window(x[g][w])) r[g][w] = ppm_ginv_damped(x[g][w], r[g][w]);
        }
#pragma unroll
        for (int g = 0; g < kGridEl; ++g)
#pragma unroll
          for (int w = 0; w < kGridNW; ++w) acc[w] = fma(el[g].x, r[g][w], acc[w]);
      }
    }
  }
#pragma unroll
  for (int w = 0; w < kGridNW; ++w) {
    const int j = jbase + w * kGridThreads;
    if (j < n_omega) out[(long long)blockIdx.z * out_split_stride + (long long)level * n_omega + j] = acc[w];
  }
}
__global__ void sigma_ppm_grid_reduce(double* __restrict__ values, const double* __restrict__ partial, long long n,
                                      int splits) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = 0.0;
    for (int s = 0; s < splits; ++s) v += partial[(long long)s * n + i];
    values[i] = v;
  }
}

// (1b) compressed grid scan.  The pole sum of one level has ~ntotal*naux poles z = e_m -/+ Omega_P but only ~1000
//     targets on a uniform grid, and 1/(w - z) is smooth in z away from w.  The pole axis is cut into bins (host plan,
//     gw.cu: ppm_grid_plan; width 0.25 Ha = the damping half-window near the targets, doubling outside the target
//     range).  Per level and bin the weights are condensed into kCmpOrder Chebyshev moments
//         mu_j = sum_{poles in bin} A T_j((z - c)/h)                                  (ppm_moments_kernel)
//     and a bin that is well separated from a chunk of 32 grid points (distance >= 3 h and >= h + 0.25, so that no pole
//     of it is inside the damping window) contributes through the Chebyshev series of the Cauchy kernel
//         sum A/(w - z) = (2 s/(h sqrt(d^2-1))) [mu_0/2 + sum_{j>=1} r^j mu_j],  d = (w-c)/h, s = sign d, r = s/(|d|+sqrt(d^2-1))
//     (truncation <= 5.83^-16 ~ 6e-13 of the bin's own contribution at the closest admissible distance, far below
//     the 1e-9 the parity tests ask for); only the poles of the remaining "near" bins (about a tenth of them) are
//     evaluated one by one, with the Rohlfing damping, exactly as in (1).  With sorted energies the poles of a bin are
//     a contiguous m-range per aux function and occupied/unoccupied segment (ppm_bin_table_kernel), so both kernels
//     stream contiguous pieces of the slab rows.  All sums run in a fixed order (deterministic).
// A warp owns a chunk of kCmpChunk = 8 consecutive grid points and walks the near poles four at a time: lane =
// (slot, point) with slot = lane / 8 the pole a lane takes out of each group of four and point = lane % 8 its grid
// point; the four slots are summed by two shuffles at the end (fixed order).  Short chunks (0.07 Ha at the default
// 0.01 Ha spacing, against the 0.5 Ha wide damping window) keep the near window small and, above all, make almost
// every group of poles warp-uniform: nobody damped (reciprocal path) or everybody damped (polynomial path); the round-1
// kernel with 32-point chunks spent most of its time in the mixed case that evaluates both.
constexpr int kCmpOrder = 16, kCmpChunk = kPpmGridChunk, kCmpSlots = 32 / kCmpChunk, kCmpWarps = 4, kCmpG = 4,
              kCmpMomentWarps = 8, kCmpMinBlocks = 5, kCmpWalk = 2;
static_assert(kCmpChunk * kCmpSlots == 32 && 32 % (kCmpSlots * kCmpG) == 0, "lane = (slot, point) mapping");

// binstart[(seg*naux + P)*(nb+1) + b] = first m of segment seg (0 occupied, 1 unoccupied) whose pole lies at or above
// edges[b]; b = 0 -> segment start, b = nb -> segment end (the outermost bins take whatever lies beyond the edges).
__global__ void ppm_bin_table_kernel(int* __restrict__ binstart, const double* __restrict__ edges, int nb,
                                     const double* __restrict__ energies, int ntotal, int n_occ,
                                     const double* __restrict__ ppm_freq, int naux) {
  const long long total = 2LL * naux * (nb + 1);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx % (nb + 1));
    const int P = (int)((idx / (nb + 1)) % naux);
    const int seg = (int)(idx / ((long long)(nb + 1) * naux));
    int lo = seg ? n_occ : 0, hi = seg ? ntotal : n_occ;
    if (b == 0) {
      hi = lo;
    } else if (b < nb) {
      const double shift = seg ? ppm_freq[P] : -ppm_freq[P];
      const double edge = edges[b];
      while (lo < hi) {                       // first m with energies[m] + shift >= edge
        const int mid = (lo + hi) >> 1;
        if (energies[mid] + shift >= edge) hi = mid; else lo = mid + 1;
      }
    }
    binstart[idx] = hi;
  }
}

// moments[((slice*n_levels + level)*nb + b)*kCmpOrder + j]: one warp per bin (lanes over the bin's m-range: coalesced),
// blockIdx.x = slice of the aux range, blockIdx.y = level; the slices are summed by sigma_ppm_grid_reduce.
__global__ void __launch_bounds__(kCmpMomentWarps * 32) ppm_moments_kernel(
    const double* __restrict__ M, long long ldn, long long slab, int naux, const double* __restrict__ energies,
    const double* __restrict__ ppm_freq, const double* __restrict__ ppm_fac, const int* __restrict__ level_slab,
    const int* __restrict__ binstart, const double* __restrict__ edges, int nb, double* __restrict__ moments) {
  const int level = blockIdx.y, n_levels = gridDim.y, slice = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double* S = M + (long long)level_slab[level] * slab;
  const int p_per = (naux + gridDim.x - 1) / gridDim.x;
  const int p_begin = slice * p_per, p_end = min(naux, p_begin + p_per);
  for (int b = warp; b < nb; b += kCmpMomentWarps) {
    const double e0 = edges[b], e1 = edges[b + 1];
    const double c = 0.5 * (e0 + e1), hinv = 2.0 / (e1 - e0);
    double mu[kCmpOrder];
#pragma unroll
    for (int j = 0; j < kCmpOrder; ++j) mu[j] = 0.0;
    // two aux functions = four (P, segment) ranges per round, so that four independent table look-ups and slab loads
    // are in flight per warp (the ranges of a core bin are only a few poles long: the loop is latency-bound)
    for (int P0 = p_begin; P0 < p_end; P0 += 2) {
      int lo[4], hi[4];
      double fac[2], shift[4];
      const double* row[2];
      int iters = 0;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int P = min(P0 + u, p_end - 1);
        fac[u] = P0 + u < p_end ? ppm_fac[P] : 0.0;
        const double Om = ppm_freq[P];
        row[u] = S + (long long)P * ldn;
#pragma unroll
        for (int seg = 0; seg < 2; ++seg) {
          const int* bs = binstart + ((long long)seg * naux + P) * (nb + 1);
          const int q = 2 * u + seg;
          lo[q] = bs[b];
          hi[q] = fac[u] != 0.0 ? bs[b + 1] : lo[q];
          shift[q] = seg ? Om : -Om;
          iters = max(iters, (hi[q] - lo[q] + 31) >> 5);
        }
      }
      for (int it = 0; it < iters; ++it) {
        double v[4], en[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int m = lo[q] + 32 * it + lane;
          const bool ok = m < hi[q];
          v[q] = ok ? row[q >> 1][m] : 0.0;
          en[q] = ok ? energies[m] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (lo[q] + 32 * it >= hi[q]) continue;            // warp-uniform: this range is exhausted
          const double a = fac[q >> 1] * v[q] * v[q];        // 0 for the lanes past the end of the range
          const double t = v[q] != 0.0 ? (en[q] + shift[q] - c) * hinv : 0.0;
          const double t2 = t + t;
          double tm = 1.0, tc = t;
          mu[0] += a;
          mu[1] = fma(a, t, mu[1]);
#pragma unroll
          for (int j = 2; j < kCmpOrder; ++j) {
            const double tn = fma(t2, tc, -tm);
            mu[j] = fma(a, tn, mu[j]);
            tm = tc;
            tc = tn;
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kCmpOrder; ++j) {
      const double v = warp_sum(mu[j]);
      if (lane == 0) moments[(((long long)slice * n_levels + level) * nb + b) * kCmpOrder + j] = v;
    }
  }
}

// Equivalent poles of a bin: for any function f that is smooth on the bin (here: the DAMPED kernel as a function of the
// pole position, for grid points whose whole damping window covers the bin), sum_i A_i f(z_i) = sum_n w_n f(z_n) with
// z_n the kCmpOrder Chebyshev-Gauss nodes of the bin and w_n = (2/K) [mu_0/2 + sum_{k>=1} mu_k T_k(t_n)]
// (Chebyshev interpolation through the nodes, integrated against the poles = the bin's moments).  The interpolation
// error of sin^2(2 pi x)/x over a bin of half-width h <= 1/16 Ha is below (2 pi h)^16 / 16! ~ 1e-20 of its maximum.
// eq[((level*nb + b)*K + n)] = (w_n, z_n): fed to the near-field loop like real poles.
__global__ void ppm_equivalent_poles_kernel(double2* __restrict__ eq, const double* __restrict__ moments,
                                            const double* __restrict__ edges, int nb, long long n_level_bins) {
  const long long total = n_level_bins * kCmpOrder;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx % kCmpOrder);
    const long long lb = idx / kCmpOrder;
    const int b = (int)(lb % nb);
    const double* mu = moments + lb * kCmpOrder;
    const double t = cospi((n + 0.5) / kCmpOrder);
    // Clenshaw for sum'_k mu_k T_k(t)
    double b1 = 0.0, b2 = 0.0;
#pragma unroll
    for (int k = kCmpOrder - 1; k >= 1; --k) {
      const double b0 = fma(2.0 * t, b1, mu[k] - b2);
      b2 = b1;
      b1 = b0;
    }
    const double sum = fma(t, b1, 0.5 * mu[0] - b2);
    const double c = 0.5 * (edges[b] + edges[b + 1]), h = 0.5 * (edges[b + 1] - edges[b]);
    eq[idx] = make_double2(sum * (2.0 / kCmpOrder), fma(h, t, c));
  }
}

// One warp per (level, chunk of kCmpChunk consecutive grid points); blockIdx.z splits the aux range of the near field
// when few warps would leave SMs idle (split 0 also adds the far field and the equivalent poles of the inner bins).
// near_range[(level*n_chunks + chunk)*4 + {0..3}]: near bins b_lo..b_hi (inclusive; lo > hi: none), of which the
// inner bins i_lo..i_hi are represented by their equivalent poles; the poles of the remaining near bins are evaluated
// one by one.
// MINB = resident CTAs per SM the register allocation is capped for (occupancy against unrolling depth; the launcher
// picks kCmpMinBlocks unless XTPB_GRID_OCC says otherwise).
// WALK = how the (up to four) near m-ranges of an aux function are turned into 32-pole tiles: 0 range by range,
// 1 as one virtual index space (short ranges share tiles), 2 like 1 with the padding lanes of the last tile carrying
// the last real pole's position at weight zero, so that a partly filled step keeps the kind (plain / damped) of its
// real poles instead of falling into the mixed branch (XTPB_GRID_WALK; measured in profiles/r02_sigma_grid_walk.jsonl).
template <int MINB, int WALK>
__global__ void __launch_bounds__(kCmpWarps * 32, MINB) sigma_ppm_grid_compressed_kernel(
    const double* __restrict__ M, long long ldn, long long slab, int naux, const double* __restrict__ energies,
    const double* __restrict__ ppm_freq, const double* __restrict__ ppm_fac, const int* __restrict__ level_slab,
    const int* __restrict__ level_mom, const double* __restrict__ omega0, double domega, int n_omega,
    const int* __restrict__ binstart,
    const double* __restrict__ edges, int nb, const int* __restrict__ near_range, int n_chunks,
    const double* __restrict__ moments, const double2* __restrict__ eq_poles, double* __restrict__ out,
    long long out_split_stride, unsigned long long* __restrict__ near_poles) {
  __shared__ double2 tile[kCmpWarps][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x * kCmpWarps + warp, level = blockIdx.y;
  if (chunk >= n_chunks) return;              // warps are independent: no block-wide barrier below
  const double* S = M + (long long)level_slab[level] * slab;
  const int mrow = level_mom ? level_mom[level] : level;      // row of the moment / equivalent-pole tables
  const int point = lane & (kCmpChunk - 1), slot = lane / kCmpChunk;
  const int j = chunk * kCmpChunk + point;
  const double om = omega0[level] + domega * (double)j;
  const int4 nr = reinterpret_cast<const int4*>(near_range)[(long long)level * n_chunks + chunk];
  const int b_lo = nr.x, b_hi = nr.y, i_lo = nr.z, i_hi = nr.w;
  const int p_per = (naux + gridDim.z - 1) / gridDim.z;
  const int p_begin = blockIdx.z * p_per, p_end = min(naux, p_begin + p_per);
  double acc4[kCmpG];                         // independent accumulation chains, summed in a fixed order
#pragma unroll
  for (int g = 0; g < kCmpG; ++g) acc4[g] = 0.0;
  long long n_near = 0;                       // poles this warp evaluates one by one (bookkeeping for the reports)

  // evaluates the 32 poles in tile[warp] (weight, position z: x = w - z; padding: weight 0, far away) against this
  // lane's grid point, kCmpSlots * kCmpG poles per step
  auto eval_tile = [&](int count) {
    constexpr int kStep = kCmpSlots * kCmpG;
    const int cnt = (count + kStep - 1) / kStep * kStep;
    for (int t = 0; t < cnt; t += kStep) {
      double2 e[kCmpG];
      double x[kCmpG], r[kCmpG];
      unsigned hmin = 0xffffffffu, hmax = 0u;
#pragma unroll
      for (int g = 0; g < kCmpG; ++g) {
        e[g] = tile[warp][t + g * kCmpSlots + slot];
        x[g] = om - e[g].y;
        const unsigned hw = static_cast<unsigned>(__double2hiint(x[g])) & 0x7fffffffu;      // |x| by its high word
        hmin = min(hmin, hw);
        hmax = max(hmax, hw);
      }
      // the poles of a range ascend and the chunk is much shorter than the damping window, so a step is almost always
      // of one kind for the whole warp: nobody damped (plain reciprocal), everybody damped (polynomial, no
      // reciprocal), or -- in two narrow zones at the edges of the window -- mixed (both, selected per element).
      // Warp-uniform branches (one warp-wide min / max of the exponent words): no divergence.
      const unsigned wmin = __reduce_min_sync(0xffffffffu, hmin), wmax = __reduce_max_sync(0xffffffffu, hmax);
      if (wmin >= 0x3fd00000u) {
#pragma unroll
        for (int g = 0; g < kCmpG; ++g) r[g] = rcp_fast(x[g]);
      } else if (wmax < 0x3fd00000u) {
#pragma unroll
        for (int g = 0; g < kCmpG; ++g) r[g] = ppm_ginv_poly(x[g]);
      } else {
#pragma unroll
        for (int g = 0; g < kCmpG; ++g) {
          const double plain = rcp_fast(x[g]), damped = ppm_ginv_poly(x[g]);
          r[g] = ppm_in_window(x[g]) ? damped : plain;
        }
      }
#pragma unroll
      for (int g = 0; g < kCmpG; ++g) acc4[g] = fma(e[g].x, r[g], acc4[g]);
    }
  };

  // ---- near field: the damped kernel, pole by pole (poles broadcast from shared memory, four per step)
  if (b_lo <= b_hi) {
    // m-ranges of aux function P that are evaluated pole by pole: per segment (occupied / unoccupied) the poles of the
    // near bins below and above the inner bins.  The table look-up of P + 1 is in flight while P is evaluated.
    int t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0, t7 = 0;
    auto ranges_of = [&](int P) {
      const int* bs0 = binstart + (long long)P * (nb + 1);
      const int* bs1 = binstart + ((long long)naux + P) * (nb + 1);
      t0 = bs0[b_lo]; t1 = bs0[i_lo]; t2 = bs0[i_hi + 1]; t3 = bs0[b_hi + 1];
      t4 = bs1[b_lo]; t5 = bs1[i_lo]; t6 = bs1[i_hi + 1]; t7 = bs1[b_hi + 1];
    };
    if (p_begin < p_end) ranges_of(p_begin);
    for (int P = p_begin; P < p_end; ++P) {
      const int c0 = t0, c1 = t1, c2 = t2, c3 = t3, c4 = t4, c5 = t5, c6 = t6, c7 = t7;
      if (P + 1 < p_end) ranges_of(P + 1);
      const double fac = ppm_fac[P];
      if (fac == 0.0) continue;
      const double Om = ppm_freq[P];
      const double* row = S + (long long)P * ldn;
      if (WALK == 0) {
#pragma unroll 1
        for (int rg = 0; rg < 4; ++rg) {
          const int lo = rg == 0 ? c0 : (rg == 1 ? c2 : (rg == 2 ? c4 : c6));
          const int hi = rg == 0 ? c1 : (rg == 1 ? c3 : (rg == 2 ? c5 : c7));
          if (lo >= hi) continue;
          const double shift = rg >= 2 ? Om : -Om;
          n_near += hi - lo;
          // raw tensor element / energy of this lane's pole in the tile that starts at m0: the arithmetic on them
          // happens when the tile is stored, so that the loads of the NEXT tile stay in flight while this one is evaluated
          double nv = 0.0, ne = 0.0;
          auto fetch = [&](int m0) {
            const int m = m0 + lane;
            nv = 0.0;
            ne = -1.0e30;
            if (m < hi) {
              nv = row[m];
              ne = energies[m];
            }
          };
          fetch(lo);
          for (int m0 = lo; m0 < hi; m0 += 32) {
            const double2 el = make_double2(fac * nv * nv, ne + shift);
            if (m0 + 32 < hi) fetch(m0 + 32);
            __syncwarp();
            tile[warp][lane] = el;
            __syncwarp();
            eval_tile(min(32, hi - m0));
          }
        }
        continue;
      }
      // The four m-ranges of this aux function (occupied / unoccupied segment, below / above the inner bins) are
      // walked as ONE virtual index space, so that short ranges (a rank of eight holds 1/8 of the levels) share tiles:
      // virtual index t -> range rg = #{offsets <= t}, m = lo_rg + (t - offset_rg); ranges 2, 3 are unoccupied (+Omega)
      const int o1 = max(c1 - c0, 0), o2 = o1 + max(c3 - c2, 0), o3 = o2 + max(c5 - c4, 0);
      const int total = o3 + max(c7 - c6, 0);
      if (total == 0) continue;
      n_near += total;
      // raw tensor element / energy of this lane's pole in the tile that starts at t0: the arithmetic on them happens
      // when the tile is stored, so that the loads of the NEXT tile stay in flight while this one is evaluated
      double nv = 0.0, ne = 0.0, nsh = 0.0;
      auto fetch = [&](int first) {
        int t = first + lane;
        nv = 0.0;
        ne = -1.0e30;
        nsh = 0.0;
        const bool real = t < total;
        if (WALK == 2) t = min(t, total - 1);      // padding lanes: position of the last real pole, weight zero
        if (t < total) {
          const int m = t < o1 ? c0 + t : (t < o2 ? c2 + (t - o1) : (t < o3 ? c4 + (t - o2) : c6 + (t - o3)));
          nv = real ? row[m] : 0.0;
          ne = energies[m];
          nsh = t < o2 ? -Om : Om;
        }
      };
      fetch(0);
      for (int v0 = 0; v0 < total; v0 += 32) {
        const double2 el = make_double2(fac * nv * nv, ne + nsh);
        if (v0 + 32 < total) fetch(v0 + 32);
        __syncwarp();
        tile[warp][lane] = el;
        __syncwarp();
        eval_tile(min(32, total - v0));
      }
    }
  }
  // ---- inner bins: their equivalent poles (all of them inside every grid point's damping window), split 0 only
  if (blockIdx.z == 0 && i_lo <= i_hi) {
    const double2* eq = eq_poles + ((long long)mrow * nb + i_lo) * kCmpOrder;
    const int total = (i_hi - i_lo + 1) * kCmpOrder;
    for (int m0 = 0; m0 < total; m0 += 32) {
      const int m = m0 + lane;
      const double2 el = m < total ? eq[m] : make_double2(0.0, -1.0e30);
      __syncwarp();
      tile[warp][lane] = el;
      __syncwarp();
      eval_tile(min(32, total - m0));
    }
  }
  double acc = 0.0;
#pragma unroll
  for (int g = 0; g < kCmpG; g += 2) acc += acc4[g] + acc4[g + 1];
  // ---- far field: Chebyshev series of the Cauchy kernel over the condensed bins, the bins dealt out to the slots
  if (blockIdx.z == 0) {
    const double* mom = moments + ((long long)mrow * nb) * kCmpOrder;
    for (int b = slot; b < nb; b += kCmpSlots) {
      if (b >= b_lo && b <= b_hi) continue;
      const double e0 = edges[b], e1 = edges[b + 1];
      const double h = 0.5 * (e1 - e0);
      const double d = (om - 0.5 * (e0 + e1)) / h;
      const double ad = fabs(d), sg = d < 0.0 ? -1.0 : 1.0;
      const double sq = sqrt(fma(ad, ad, -1.0));
      const double r = sg / (ad + sq);
      const double* mu = mom + (long long)b * kCmpOrder;
      double f = 0.0;
#pragma unroll
      for (int jj = kCmpOrder - 1; jj >= 1; --jj) f = (f + mu[jj]) * r;
      f = fma(0.5, mu[0], f);
      acc = fma(f, 2.0 * sg / (sq * h), acc);
    }
  }
  // sum of the slots: lanes point, point + 8, point + 16, point + 24 (fixed order)
#pragma unroll
  for (int o = kCmpChunk; o < 32; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (slot == 0 && j < n_omega) out[(long long)blockIdx.z * out_split_stride + (long long)level * n_omega + j] = acc;
  if (lane == 0 && n_near > 0) atomicAdd(near_poles, (unsigned long long)n_near);   // integer: order-independent
}

// (1c) CTA-cooperative near field.  The four warps of a CTA own four consecutive chunks of ONE level, so their near
//     ranges overlap almost completely; in (1b) every warp nevertheless looks up its own eight table entries per aux
//     function (latency-bound global loads) and stages its own 32-pole tiles -- at C60 size about as many instructions
//     as the evaluations themselves (ncu: 59 thread instructions per evaluated pair against ~19 of arithmetic).  Here
//     the CTA works through the aux functions together:
//       * the table entries of bins B_lo .. B_hi + 1 (union of the warps' near bins) of aux function P + 1 and the
//         poles of the union range of P + 1 (raw tensor elements and energies, up to kCtaCap of them) are loaded by all
//         128 threads, coalesced, into registers while P is being evaluated, and stored to shared memory (double-
//         buffered) at the top of the next round: ONE barrier per aux function;
//       * every warp reads its eight range bounds from shared memory and evaluates its own sub-ranges of the staged
//         poles straight from the shared pole array (same step structure as (1b), padding lanes carry the last real
//         pole at weight zero);
//       * union ranges longer than kCtaCap are finished in further pieces, loaded synchronously (rare).
//     Inner bins (equivalent poles), far field, reduction and output are those of (1b).
constexpr int kCtaCap = 1024, kCtaPre = kCtaCap / (kCmpWarps * 32), kCtaMaxTab = 64, kCtaMinBlocks = 4;

template <int MINB>
__global__ void __launch_bounds__(kCmpWarps * 32, MINB) sigma_ppm_grid_cta_kernel(
    const double* __restrict__ M, long long ldn, long long slab, int naux, const double* __restrict__ energies,
    const double* __restrict__ ppm_freq, const double* __restrict__ ppm_fac, const int* __restrict__ level_slab,
    const int* __restrict__ level_mom, const double* __restrict__ omega0, double domega, int n_omega,
    const int* __restrict__ binstart,
    const double* __restrict__ edges, int nb, const int* __restrict__ near_range, int n_chunks,
    const double* __restrict__ moments, const double2* __restrict__ eq_poles, double* __restrict__ out,
    long long out_split_stride, unsigned long long* __restrict__ near_poles) {
  constexpr int kThreads = kCmpWarps * 32;
  __shared__ double2 poles[2][kCtaCap];
  __shared__ double2 tile[kCmpWarps][32];
  __shared__ int tab[2][2][kCtaMaxTab];
  __shared__ int cta_lo, cta_hi;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunk = blockIdx.x * kCmpWarps + warp, level = blockIdx.y;
  const bool active = chunk < n_chunks;       // idle warps of the last CTA of a level still load and synchronise
  const double* S = M + (long long)level_slab[level] * slab;
  const int mrow = level_mom ? level_mom[level] : level;
  const int point = lane & (kCmpChunk - 1), slot = lane / kCmpChunk;
  const int j = chunk * kCmpChunk + point;
  const double om = omega0[level] + domega * (double)j;
  int b_lo = 1, b_hi = 0, i_lo = 1, i_hi = 0;
  if (active) {
    const int4 nr = reinterpret_cast<const int4*>(near_range)[(long long)level * n_chunks + chunk];
    b_lo = nr.x; b_hi = nr.y; i_lo = nr.z; i_hi = nr.w;
  }
  const bool has_near = b_lo <= b_hi;
  if (tid == 0) { cta_lo = 0x7fffffff; cta_hi = -1; }
  __syncthreads();
  if (lane == 0 && has_near) { atomicMin(&cta_lo, b_lo); atomicMax(&cta_hi, b_hi); }
  __syncthreads();
  const int B_lo = cta_lo, B_hi = cta_hi;
  const int ntab = B_hi - B_lo + 2;           // table entries b = B_lo .. B_hi + 1 (<= kCtaMaxTab, checked by the host)
  const int p_per = (naux + gridDim.z - 1) / gridDim.z;
  const int p_begin = blockIdx.z * p_per, p_end = min(naux, p_begin + p_per);
  double acc4[kCmpG];
#pragma unroll
  for (int g = 0; g < kCmpG; ++g) acc4[g] = 0.0;
  long long n_near = 0;

  // evaluates src[a .. b) (weight, position) against this lane's grid point, kCmpSlots * kCmpG poles per step; the
  // lanes past the end of the range read its last pole at weight zero (keeps the step's kind)
  auto eval_range = [&](const double2* src, int a, int b) {
    constexpr int kStep = kCmpSlots * kCmpG;
    const int last = b - 1;
    for (int t = a; t < b; t += kStep) {
      double2 e[kCmpG];
      double x[kCmpG], r[kCmpG];
      unsigned hmin = 0xffffffffu, hmax = 0u;
#pragma unroll
      for (int g = 0; g < kCmpG; ++g) {
        const int idx = t + g * kCmpSlots + slot;
        e[g] = src[min(idx, last)];
        if (idx > last) e[g].x = 0.0;
        x[g] = om - e[g].y;
        const unsigned hw = static_cast<unsigned>(__double2hiint(x[g])) & 0x7fffffffu;
        hmin = min(hmin, hw);
        hmax = max(hmax, hw);
      }
      const unsigned wmin = __reduce_min_sync(0xffffffffu, hmin), wmax = __reduce_max_sync(0xffffffffu, hmax);
      if (wmin >= 0x3fd00000u) {
#pragma unroll
        for (int g = 0; g < kCmpG; ++g) r[g] = rcp_fast(x[g]);
      } else if (wmax < 0x3fd00000u) {
#pragma unroll
        for (int g = 0; g < kCmpG; ++g) r[g] = ppm_ginv_poly(x[g]);
      } else {
#pragma unroll
        for (int g = 0; g < kCmpG; ++g) {
          const double plain = rcp_fast(x[g]), damped = ppm_ginv_poly(x[g]);
          r[g] = ppm_in_window(x[g]) ? damped : plain;
        }
      }
#pragma unroll
      for (int g = 0; g < kCmpG; ++g) acc4[g] = fma(e[g].x, r[g], acc4[g]);
    }
  };

  if (B_lo <= B_hi && p_begin < p_end) {
    // state of an aux function: union range [u0, u0 + nA) of the occupied segment (poles e - Omega) followed by
    // [u2, u2 + nB) of the unoccupied one (e + Omega) as one index space t; this warp's four sub-ranges in m
    struct PState { int u0, u2, nA, total; double fac, Om; int c[8]; };
    auto read_state = [&](int tb, int P, PState& st) {
      const int* t0 = tab[tb][0];
      const int* t1 = tab[tb][1];
      st.fac = ppm_fac[P];
      st.Om = ppm_freq[P];
      st.u0 = t0[0];
      st.u2 = t1[0];
      st.nA = max(t0[ntab - 1] - st.u0, 0);
      st.total = st.fac != 0.0 ? st.nA + max(t1[ntab - 1] - st.u2, 0) : 0;
      if (has_near) {
        st.c[0] = t0[b_lo - B_lo]; st.c[1] = t0[i_lo - B_lo]; st.c[2] = t0[i_hi + 1 - B_lo]; st.c[3] = t0[b_hi + 1 - B_lo];
        st.c[4] = t1[b_lo - B_lo]; st.c[5] = t1[i_lo - B_lo]; st.c[6] = t1[i_hi + 1 - B_lo]; st.c[7] = t1[b_hi + 1 - B_lo];
      }
    };
    auto tab_entry = [&](int P) -> int {        // this thread's entry of the table rows of P (tid < 2 * ntab)
      const int seg = tid / ntab, b = B_lo + tid % ntab;
      return binstart[((long long)seg * naux + P) * (nb + 1) + b];
    };
    double pv[kCtaPre], pe[kCtaPre];
    // raw loads of piece [ps, ps + kCtaCap) of the aux function with state st, element q * kThreads + tid per thread
    auto load_piece = [&](const PState& st, int P, int ps) {
      const double* row = S + (long long)P * ldn;
#pragma unroll
      for (int q = 0; q < kCtaPre; ++q) {
        const int t = ps + q * kThreads + tid;
        pv[q] = 0.0;
        pe[q] = 0.0;
        if (t < st.total) {
          const int m = t < st.nA ? st.u0 + t : st.u2 + (t - st.nA);
          pv[q] = row[m];
          pe[q] = energies[m];
        }
      }
    };
    auto store_piece = [&](const PState& st, int buf, int ps) {
#pragma unroll
      for (int q = 0; q < kCtaPre; ++q) {
        const int t = ps + q * kThreads + tid;
        if (t < st.total)
          poles[buf][t - ps] = make_double2(st.fac * pv[q] * pv[q], pe[q] + (t < st.nA ? -st.Om : st.Om));
      }
    };
    // this warp's sub-ranges of piece [ps, pe) of the staged poles
    auto eval_piece = [&](const PState& st, int buf, int ps, int pend) {
      if (!has_near) return;
#pragma unroll
      for (int rg = 0; rg < 4; ++rg) {
        const int lo = st.c[2 * rg], hi = st.c[2 * rg + 1];
        if (lo >= hi) continue;
        const int off = rg < 2 ? -st.u0 : st.nA - st.u2;      // m -> t
        const int ta = max(lo + off, ps), tb = min(hi + off, pend);
        if (ta < tb) eval_range(poles[buf], ta - ps, tb - ps);
      }
    };

    PState cur, nxt;
    int tab_reg = 0;
    if (tid < 2 * ntab) tab[0][tid / ntab][tid % ntab] = tab_entry(p_begin);
    __syncthreads();
    read_state(0, p_begin, nxt);
    load_piece(nxt, p_begin, 0);
    if (p_begin + 1 < p_end && tid < 2 * ntab) tab_reg = tab_entry(p_begin + 1);
    for (int P = p_begin; P < p_end; ++P) {
      const int buf = (P - p_begin) & 1;
      cur = nxt;
      store_piece(cur, buf, 0);
      if (P + 1 < p_end && tid < 2 * ntab) tab[buf ^ 1][tid / ntab][tid % ntab] = tab_reg;
      __syncthreads();
      if (P + 1 < p_end) {
        read_state(buf ^ 1, P + 1, nxt);
        load_piece(nxt, P + 1, 0);
        if (P + 2 < p_end && tid < 2 * ntab) tab_reg = tab_entry(P + 2);
      }
      if (cur.total == 0) continue;
      if (has_near)
        n_near += max(cur.c[1] - cur.c[0], 0) + max(cur.c[3] - cur.c[2], 0) + max(cur.c[5] - cur.c[4], 0) +
                  max(cur.c[7] - cur.c[6], 0);
      eval_piece(cur, buf, 0, min(cur.total, kCtaCap));
      if (cur.total > kCtaCap) {
        // the rest of a long union range, piece by piece through the same buffer; the prefetched registers of P + 1
        // are parked in shared memory-free fashion: they are simply reloaded afterwards
        for (int ps = kCtaCap; ps < cur.total; ps += kCtaCap) {
          __syncthreads();                     // everybody is done with the previous piece
          load_piece(cur, P, ps);
          store_piece(cur, buf, ps);
          __syncthreads();
          eval_piece(cur, buf, ps, min(cur.total, ps + kCtaCap));
        }
        if (P + 1 < p_end) load_piece(nxt, P + 1, 0);
      }
    }
  }
  // ---- inner bins: their equivalent poles, split 0 only
  if (active && blockIdx.z == 0 && i_lo <= i_hi) {
    const double2* eq = eq_poles + ((long long)mrow * nb + i_lo) * kCmpOrder;
    const int total = (i_hi - i_lo + 1) * kCmpOrder;
    for (int m0 = 0; m0 < total; m0 += 32) {
      const int m = m0 + lane;
      const double2 el = m < total ? eq[m] : make_double2(0.0, -1.0e30);
      __syncwarp();
      tile[warp][lane] = el;
      __syncwarp();
      eval_range(tile[warp], 0, min(32, total - m0));
    }
  }
  if (!active) return;
  double acc = 0.0;
#pragma unroll
  for (int g = 0; g < kCmpG; g += 2) acc += acc4[g] + acc4[g + 1];
  // ---- far field: Chebyshev series of the Cauchy kernel over the condensed bins, the bins dealt out to the slots
  if (blockIdx.z == 0) {
    const double* mom = moments + ((long long)mrow * nb) * kCmpOrder;
    for (int b = slot; b < nb; b += kCmpSlots) {
      if (b >= b_lo && b <= b_hi) continue;
      const double e0 = edges[b], e1 = edges[b + 1];
      const double h = 0.5 * (e1 - e0);
      const double d = (om - 0.5 * (e0 + e1)) / h;
      const double ad = fabs(d), sg = d < 0.0 ? -1.0 : 1.0;
      const double sq = sqrt(fma(ad, ad, -1.0));
      const double r = sg / (ad + sq);
      const double* mu = mom + (long long)b * kCmpOrder;
      double f = 0.0;
#pragma unroll
      for (int jj = kCmpOrder - 1; jj >= 1; --jj) f = (f + mu[jj]) * r;
      f = fma(0.5, mu[0], f);
      acc = fma(f, 2.0 * sg / (sq * h), acc);
    }
  }
#pragma unroll
  for (int o = kCmpChunk; o < 32; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (slot == 0 && j < n_omega) out[(long long)blockIdx.z * out_split_stride + (long long)level * n_omega + j] = acc;
  if (lane == 0 && n_near > 0) atomicAdd(near_poles, (unsigned long long)n_near);
}

// (2) pair kernel: arbitrary (level, frequency) pairs (bisection steps, final Sigma_c, derivatives).
//     One CTA column per pair, kPairChunks CTAs split the aux range; deterministic two-stage reduction.
//     Bound: HBM/L2 (each pair streams its slab once).
constexpr int kPairThreads = 256, kPairChunks = 32;

__global__ void __launch_bounds__(kPairThreads) sigma_ppm_pairs_kernel(
    const double* __restrict__ M, long long ldn, long long slab, int ntotal, int naux, int n_occ,
    const double* __restrict__ energies, const double* __restrict__ ppm_freq, const double* __restrict__ ppm_fac,
    const int* __restrict__ pair_slab, const double* __restrict__ pair_omega, double* __restrict__ partial) {
  __shared__ double sh[kPairThreads / 32];
  const int pair = blockIdx.y, chunk = blockIdx.x;
  const double* S = M + (long long)pair_slab[pair] * slab;
  const double om = pair_omega[pair];
  const int per = (naux + kPairChunks - 1) / kPairChunks;
  const int p_begin = chunk * per, p_end = min(naux, p_begin + per);
  double val = 0.0, der = 0.0;
  for (int P = p_begin; P < p_end; ++P) {
    const double fac = ppm_fac[P];
    if (fac == 0.0) continue;
    const double Om = ppm_freq[P];
    const double* row = S + (long long)P * ldn;
    double v1 = 0.0, d1 = 0.0;
    for (int m = threadIdx.x; m < ntotal; m += kPairThreads) {
      const double v = row[m];
      const double g = ppm_ginv(om - energies[m] + (m < n_occ ? Om : -Om));
      const double a = v * v * g;
      v1 += a;
      d1 -= a * g;
    }
    val += fac * v1;
    der += fac * d1;
  }
  val = block_sum<kPairThreads>(val, sh);
  der = block_sum<kPairThreads>(der, sh);
  if (threadIdx.x == 0) {
    partial[((long long)pair * kPairChunks + chunk) * 2 + 0] = val;
    partial[((long long)pair * kPairChunks + chunk) * 2 + 1] = der;
  }
}
__global__ void sigma_ppm_pairs_finalize(const double* __restrict__ partial, int n_pairs, double* values,
                                         double* derivs) {
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= n_pairs) return;
  double v = 0.0, d = 0.0;
  for (int c = 0; c < kPairChunks; ++c) {
    v += partial[((long long)pair * kPairChunks + c) * 2 + 0];
    d += partial[((long long)pair * kPairChunks + c) * 2 + 1];
  }
  values[pair] = v;
  if (derivs) derivs[pair] = d;
}

// (3) off-diagonal elements as a contraction (upstream Sigma_PPM::CalcCorrelationOffDiagElement does it pair by
//     pair): W[l][p][m] = fac_P ginv(w_l - z_Pm) M~[l][P][m]; then S = W * M~^T and Sigma_c = (S + S^T)/2.
__global__ void sigma_ppm_weighted_slab_kernel(double* __restrict__ W, const double* __restrict__ M, long long ldn,
                                               long long slab, int ntotal, int p0, int pcnt, int n_occ,
                                               const double* __restrict__ energies,
                                               const double* __restrict__ ppm_freq,
                                               const double* __restrict__ ppm_fac, int slab0, int n_levels,
                                               const double* __restrict__ level_omega) {
  const long long per_level = (long long)pcnt * ldn;
  const long long total = per_level * n_levels;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(idx % ldn);
    const int p = (int)((idx / ldn) % pcnt);
    const int l = (int)(idx / per_level);
    double out = 0.0;
    if (m < ntotal) {
      const int P = p0 + p;
      const double fac = ppm_fac[P];
      if (fac != 0.0) {
        const double Om = ppm_freq[P];
        const double g = ppm_ginv(level_omega[l] - energies[m] + (m < n_occ ? Om : -Om));
        out = fac * g * M[(long long)(slab0 + l) * slab + (long long)P * ldn + m];
      }
    }
    W[idx] = out;
  }
}

// d[i] += alpha sum_p F[p*ld + i]^2: threads over i (coalesced), serial over the short p range
__global__ void add_column_square_sums_kernel(double* __restrict__ d, const double* __restrict__ F, long long ld,
                                              int rows, long long n, double alpha) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int p = 0; p < rows; ++p) {
      const double v = F[(long long)p * ld + i];
      acc = fma(v, v, acc);
    }
    d[i] = fma(alpha, acc, d[i]);
  }
}

// ------------------------------------------------------------------ BSE diagonal
// Upstream BSE_OPERATOR::diagonal (bse_operator.cc).  Block = one valence index v, threads = conduction index c.
__global__ void bse_diag_helpers_kernel(double* __restrict__ dvv, double* __restrict__ dcc, int vt, int ct, int naux,
                                        const double* __restrict__ Mvv, long long ldvv, long long slabvv,
                                        const double* __restrict__ Mcc, long long ldcc, long long slabcc) {
  const long long nv = (long long)vt * naux, nc = (long long)ct * naux;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < nv + nc;
       idx += (long long)gridDim.x * blockDim.x) {
    if (idx < nv) {
      const int P = (int)(idx % naux), v = (int)(idx / naux);
      dvv[idx] = Mvv[(long long)v * slabvv + (long long)P * ldvv + v];          // carries eps_inv[P]
    } else {
      const long long j = idx - nv;
      const int c = (int)(j % ct), P = (int)(j / ct);
      dcc[j] = Mcc[(long long)c * slabcc + (long long)P * ldcc + c];            // dcc[P][c]
    }
  }
}
__global__ void bse_diagonal_kernel(double* __restrict__ diag, int vt, int ct, int naux, const double* __restrict__ Mvc,
                                    long long ldvc, long long slabvc, const double* __restrict__ dvv,
                                    const double* __restrict__ dcc, const double* __restrict__ Mcv, long long ldcv,
                                    long long slabcv, const double* __restrict__ hqp_diag, int cqp, int cx, int cd,
                                    int cd2) {
  const int v = blockIdx.x;
  for (int c = threadIdx.x; c < ct; c += blockDim.x) {
    double acc = 0.0;
    if (cqp) acc += cqp * (hqp_diag[vt + c] - hqp_diag[v]);
    if (cx || cd || cd2) {
      double sx = 0.0, sd = 0.0, sd2 = 0.0;
      for (int P = 0; P < naux; ++P) {
        const double mvc = (cx || cd2) ? Mvc[(long long)v * slabvc + (long long)P * ldvc + c] : 0.0;
        if (cx) sx += mvc * mvc;
        if (cd) sd += dvv[(long long)v * naux + P] * dcc[(long long)P * ct + c];
        if (cd2) sd2 += Mcv[(long long)c * slabcv + (long long)P * ldcv + v] * mvc;   // Mcv carries eps_inv[P]
      }
      acc += cx * sx - cd * sd - cd2 * sd2;
    }
    diag[(long long)v * ct + c] = acc;
  }
}

// ------------------------------------------------------------------ Davidson vector kernels
__global__ void __launch_bounds__(256) col_norms_kernel(const double* __restrict__ A, long long ld, long long rows,
                                                        double* out) {
  __shared__ double sh[8];
  const double* col = A + (long long)blockIdx.x * ld;
  double s = 0.0;
  for (long long i = threadIdx.x; i < rows; i += 256) s += col[i] * col[i];
  s = block_sum<256>(s, sh);
  if (threadIdx.x == 0) out[blockIdx.x] = sqrt(s);
}
__global__ void __launch_bounds__(256) column_dots_kernel(const double* __restrict__ A, long long lda,
                                                          const double* __restrict__ B, long long ldb, long long rows,
                                                          double* out) {
  __shared__ double sh[8];
  const double* a = A + (long long)blockIdx.x * lda;
  const double* b = B + (long long)blockIdx.x * ldb;
  double s = 0.0;
  for (long long i = threadIdx.x; i < rows; i += 256) s += a[i] * b[i];
  s = block_sum<256>(s, sh);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}
__global__ void residuals_kernel(double* res, long long ldr, const double* __restrict__ q, long long ldq,
                                 const double* __restrict__ lambda, long long rows, int cols) {
  const long long total = rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx % rows;
    const int c = (int)(idx / rows);
    res[r + c * ldr] -= lambda[c] * q[r + c * ldq];
  }
}
// t = r / (lambda - D); xd = x / (lambda - D); partial dots (x.t, x.xd) per block -> scratch
__global__ void __launch_bounds__(256) correction_kernel(double* __restrict__ t, double* __restrict__ xd,
                                                         const double* __restrict__ r, const double* __restrict__ x,
                                                         const double* __restrict__ D, double lambda, long long n,
                                                         int olsen, double* __restrict__ partial) {
  __shared__ double sh[8];
  double d1 = 0.0, d2 = 0.0;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += 256LL * gridDim.x) {
    double den = lambda - D[i];
    if (fabs(den) < 1e-12) den = 1e-12;
    const double ti = r[i] / den;
    t[i] = ti;
    if (olsen) {
      const double xi = x[i], xdi = xi / den;
      xd[i] = xdi;
      d1 += xi * ti;
      d2 += xi * xdi;
    }
  }
  if (olsen) {
    d1 = block_sum<256>(d1, sh);
    d2 = block_sum<256>(d2, sh);
    if (threadIdx.x == 0) {
      partial[2 * blockIdx.x] = d1;
      partial[2 * blockIdx.x + 1] = d2;
    }
  }
}
__global__ void __launch_bounds__(256) olsen_finish_kernel(double* __restrict__ t, const double* __restrict__ xd,
                                                           const double* __restrict__ partial, int nblocks,
                                                           long long n) {
  __shared__ double eps_sh;
  if (threadIdx.x == 0) {
    double d1 = 0.0, d2 = 0.0;
    for (int b = 0; b < nblocks; ++b) {
      d1 += partial[2 * b];
      d2 += partial[2 * b + 1];
    }
    eps_sh = d1 / d2;
  }
  __syncthreads();
  const double eps = eps_sh;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += 256LL * gridDim.x) t[i] -= eps * xd[i];
}
__global__ void unit_vectors_kernel(double* V, long long ld, const long long* __restrict__ idx, int cols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < cols) V[idx[c] + (long long)c * ld] = 1.0;
}


// ------------------------------------------------------------------ multi-GPU index maps (cyclic second index)
__global__ void gather_strided_kernel(double* __restrict__ dst, const double* __restrict__ src, long long first,
                                      long long stride, long long n) {
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x)
    dst[j] = src[first + j * stride];
}
template <bool TO_LOCAL>
__global__ void cyclic_cols_kernel(double* __restrict__ dst, long long ld_dst, const double* __restrict__ src,
                                   long long ld_src, long long rows, long long ncols_loc, int rank, int world) {
  const long long total = rows * ncols_loc;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long jl = idx % ncols_loc, r = idx / ncols_loc;
    const long long g = rank + jl * world;
    if (TO_LOCAL) dst[r * ld_dst + jl] = src[r * ld_src + g];
    else dst[r * ld_dst + g] = src[r * ld_src + jl];
  }
}
// dst[i][Ql][j] <- G[s][i][P0+Ql][jl], global column g = n0 + j held by rank s = g % world at local window
// position jl = g / world - (first local column of s inside the window)
__global__ void window_from_gathered_kernel(double* __restrict__ dst, long long dst_ld, long long dst_slab,
                                            const double* __restrict__ G, long long ldl, int mcnt, int naux, int P0,
                                            int pcnt, int n0, int ncnt, int world) {
  const long long total = (long long)mcnt * pcnt * ncnt;
  const long long block = (long long)mcnt * naux * ldl;      // one rank's gathered block
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % ncnt);
    const int Ql = (int)((idx / ncnt) % pcnt);
    const int i = (int)(idx / ((long long)ncnt * pcnt));
    const int g = n0 + j;
    const int s = g % world;
    const int first = n0 > s ? (n0 - s + world - 1) / world : 0;
    const int jl = g / world - first;
    dst[(long long)i * dst_slab + (long long)Ql * dst_ld + j] =
        G[(long long)s * block + ((long long)i * naux + P0 + Ql) * ldl + jl];
  }
}


// ------------------------------------------------------------------ dense BSE Hamiltonian: Hqp part
// part 0: rows (v2, c1) of column (v2l, c2) += cqp * Hqp[vt+c1, vt+c2];  part 1: rows (v1, c2) -= cqp * Hqp[v1, v2]
__global__ void bse_add_hqp_kernel(double* __restrict__ H, long long ld, int vt, int ct, int v2lo, int ns,
                                   const double* __restrict__ hqp, long long hs, double cqp, int part) {
  const int per = part == 0 ? ct : vt;
  const long long total = (long long)ns * ct * per;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(idx % per);
    const long long col = idx / per;
    const int c2 = (int)(col % ct), v2 = v2lo + (int)(col / ct);
    if (part == 0) H[(long long)v2 * ct + t + col * ld] += cqp * hqp[(vt + t) + (long long)(vt + c2) * hs];
    else H[(long long)t * ct + c2 + col * ld] -= cqp * hqp[t + (long long)v2 * hs];
  }
}

// ------------------------------------------------------------------ dense BSE Hamiltonian: (v1,c1) <-> (v2,c2) symmetry
// The screened direct term Hd[(v1,c1),(v2,c2)] = sum_P Mcc[c2][P][c1] w_P Mvv[v2][P][v1] is symmetric under the
// simultaneous swap v1 <-> v2, c1 <-> c2, so only the occupied pairs v1 <= v2 are contracted (half the flops).
// t = v2 (v2 + 1) / 2 + v1 enumerates them.
__device__ __forceinline__ void pair_of_index(long long t, int& v2, int& v1) {
  int a = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((long long)(a + 1) * (a + 2) / 2 <= t) ++a;
  while ((long long)a * (a + 1) / 2 > t) --a;
  v2 = a;
  v1 = (int)(t - (long long)a * (a + 1) / 2);
}
// Ftri[P*ldt + t] = F[P*ldf + v2*vt + v1]: the pair columns of the flat operand [P][v2][v1], packed
__global__ void bse_pack_pairs_kernel(double* __restrict__ Ftri, long long ldt, const double* __restrict__ F,
                                      long long ldf, int vt, long long npairs, int naux) {
  const long long total = npairs * naux;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long t = idx % npairs;
    const int P = (int)(idx / npairs);
    int v2, v1;
    pair_of_index(t, v2, v1);
    Ftri[(long long)P * ldt + t] = F[(long long)P * ldf + (long long)v2 * vt + v1];
  }
}
// T[(c1 + c2*ct) + tl*ct*ct], tl < tcnt (pairs t0 + tl), holds alpha * Hd for the pair's ct x ct block:
//   H[(v1*ct + c1) + (v2*ct + c2)*ld] += T   and, for v1 != v2, the mirrored H[(v2*ct + c2) + (v1*ct + c1)*ld] += T
// (a 32 x 32 tile per block; the mirror goes through a shared-memory transpose so that both writes are coalesced).
__global__ void __launch_bounds__(256) bse_scatter_pairs_kernel(double* __restrict__ H, long long ld, int ct,
                                                                const double* __restrict__ T, long long t0) {
  __shared__ double tile[32][33];
  const int tiles = (ct + 31) / 32;
  const int tc1 = blockIdx.x % tiles, tc2 = blockIdx.x / tiles;
  const long long tl = blockIdx.y;
  int v2, v1;
  pair_of_index(t0 + tl, v2, v1);
  const double* src = T + tl * (long long)ct * ct;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int c1 = tc1 * 32 + tx, c2 = tc2 * 32 + r;
    double v = 0.0;
    if (c1 < ct && c2 < ct) {
      v = src[c1 + (long long)c2 * ct];
      H[((long long)v1 * ct + c1) + ((long long)v2 * ct + c2) * ld] += v;
    }
    tile[r][tx] = v;
  }
  if (v1 == v2) return;
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c2 = tc2 * 32 + tx, c1 = tc1 * 32 + r;
    if (c1 < ct && c2 < ct) H[((long long)v2 * ct + c2) + ((long long)v1 * ct + c1) * ld] += tile[tx][r];
  }
}
}  // namespace

// ------------------------------------------------------------------ host wrappers
void k_chi0_weights(double* d, const double* e_m, const double* e_n, int n_occ, int n_occ_n, int a0, int K,
                    const double* omegas_dev, int n_omega, bool imag, double eta, cudaStream_t s) {
  const long long total = (long long)n_omega * n_occ * K;
  chi0_weights_kernel<<<blocks_for(total, 256, 4096), 256, 0, s>>>(d, e_m, e_n, n_occ, n_occ_n, a0, K, omegas_dev,
                                                                   n_omega, imag ? 1 : 0, eta);
  LAUNCH_CHECK();
}
void k_set_identity(double* A, int n, long long ld, cudaStream_t s) {
  set_identity_kernel<<<blocks_for((long long)n * n, 256, 4096), 256, 0, s>>>(A, n, ld);
  LAUNCH_CHECK();
}
void k_add_diagonal(double* A, int n, long long ld, double v, cudaStream_t s) {
  add_diagonal_kernel<<<blocks_for(n, 256), 256, 0, s>>>(A, n, ld, v);
  LAUNCH_CHECK();
}
void k_scale_columns(double* A, int rows, int cols, long long ld, const double* scale, cudaStream_t s) {
  scale_columns_kernel<<<blocks_for((long long)rows * cols, 256, 4096), 256, 0, s>>>(A, rows, cols, ld, scale);
  LAUNCH_CHECK();
}
void k_zero_strict_lower(double* A, int n, long long ld, cudaStream_t s) {
  zero_strict_lower_kernel<<<blocks_for((long long)n * n, 256, 8192), 256, 0, s>>>(A, n, ld);
  LAUNCH_CHECK();
}
void k_extract_diagonal(const double* A, int n, long long ld, double* out, cudaStream_t s) {
  extract_diagonal_kernel<<<blocks_for(n, 256), 256, 0, s>>>(A, n, ld, out);
  LAUNCH_CHECK();
}
void k_copy_2d(double* dst, long long ldd, const double* src, long long lds, int rows, long long cols, cudaStream_t s) {
  copy_2d_kernel<<<blocks_for((long long)rows * cols, 256, 8192), 256, 0, s>>>(dst, ldd, src, lds, rows, cols);
  LAUNCH_CHECK();
}
void k_scale(double* x, long long n, double a, cudaStream_t s) {
  scale_kernel<<<blocks_for(n, 256, 4096), 256, 0, s>>>(x, n, a);
  LAUNCH_CHECK();
}
void k_axpby(double* y, const double* x, long long n, double a, double b, cudaStream_t s) {
  axpby_kernel<<<blocks_for(n, 256, 4096), 256, 0, s>>>(y, x, n, a, b);
  LAUNCH_CHECK();
}
void k_extract_window(double* dst, long long dst_ld, long long dst_slab, const double* M, long long ldn, long long slab,
                      int m0, int mcnt, int n0, int ncnt, int naux, const double* scale, cudaStream_t s) {
  const long long total = (long long)mcnt * naux * ncnt;
  extract_window_kernel<<<blocks_for(total, 256, 8192), 256, 0, s>>>(dst, dst_ld, dst_slab, M, ldn, slab, m0, mcnt, n0,
                                                                    ncnt, naux, scale);
  LAUNCH_CHECK();
}

void k_add_column_square_sums(double* d, const double* F, long long ld, int rows, long long n, double alpha,
                              cudaStream_t s) {
  if (n <= 0) return;
  add_column_square_sums_kernel<<<blocks_for(n, 128, 4096), 128, 0, s>>>(d, F, ld, rows, n, alpha);
  LAUNCH_CHECK();
}
void k_sigma_ppm_grid(const double* M, long long ldn, long long slab, int ntotal, int naux, int n_occ,
                      const double* energies, const double* ppm_freq, const double* ppm_fac, const int* level_slab,
                      const double* omega0, double domega, int n_omega, int n_levels, double* values, cudaStream_t s) {
  const int per_block = kGridThreads * kGridNW;
  const int bx = (n_omega + per_block - 1) / per_block;
  // ~8 resident CTAs per SM: split the aux range when there are few (level, frequency-chunk) blocks
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int splits = (8 * sms + bx * n_levels - 1) / (bx * n_levels);
  splits = std::min(splits, 16);
  splits = std::min(splits, naux / 64 > 0 ? naux / 64 : 1);
  splits = std::max(splits, 1);
  dim3 grid(bx, n_levels, splits);
  const long long n = (long long)n_levels * n_omega;
  DBuf partial;
  if (splits > 1) partial.alloc((size_t)(n * splits));
  // work = pole evaluations (one reciprocal + 5 FP64 operations each)
  const int slot = prof_begin(PROF_SIGMA_GRID, (double)ntotal * naux * (double)n_omega * n_levels, s);
  sigma_ppm_grid_kernel<<<grid, kGridThreads, 0, s>>>(M, ldn, slab, ntotal, naux, n_occ, energies, ppm_freq, ppm_fac,
                                                      level_slab, omega0, domega, n_omega,
                                                      splits > 1 ? partial.p : values, n);
  LAUNCH_CHECK();
  if (splits > 1) {
    sigma_ppm_grid_reduce<<<blocks_for(n, 256, 2048), 256, 0, s>>>(values, partial.p, n, splits);
    LAUNCH_CHECK();
  }
  prof_end(slot, s);
  if (splits > 1) XTPB_CUDA(cudaStreamSynchronize(s));   // partial is freed on return
}
// Compressed scan, part 1: everything that depends on the tensor, the energies and the plasmon-pole parameters but not
// on the target frequencies -- the bin table, the Chebyshev moments per (level, bin) (one stream over the slabs) and
// the equivalent poles.  The state outlives the call: GW::grid_scan builds it, the bisection rounds and the final
// Sigma_c of GW::SolveQP evaluate further points through it (k_ppm_scan_evaluate) without touching the slabs again.
void k_ppm_scan_prepare(const double* M, long long ldn, long long slab, int ntotal, int naux, int n_occ,
                        const double* energies, const double* ppm_freq, const double* ppm_fac, const int* level_slab,
                        int n_levels, const double* edges_host, int nb, PpmScanState& st, cudaStream_t s) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int slices = (4 * sms + n_levels - 1) / n_levels;
  slices = std::max(1, std::min(slices, std::min(16, std::max(1, naux / 64))));
  const long long n_mom = (long long)n_levels * nb * kCmpOrder;
  const long long n_table = 2LL * naux * (nb + 1);
  st.valid = false;
  st.nb = nb;
  st.n_levels = n_levels;
  st.edges.assign(edges_host, edges_host + nb + 1);
  st.edges_dev.ensure((size_t)(nb + 1));
  st.table.ensure((size_t)((n_table + 1) / 2));
  st.eq.ensure((size_t)(2 * n_mom));            // equivalent poles: (weight, position) per (level, bin, node)
  st.mom.ensure((size_t)n_mom);
  DBuf mom_part;
  if (slices > 1) mom_part.alloc((size_t)(n_mom * slices));
  int* table_i = reinterpret_cast<int*>(st.table.p);
  XTPB_CUDA(cudaMemcpyAsync(st.edges_dev.p, st.edges.data(), (size_t)(nb + 1) * sizeof(double), cudaMemcpyHostToDevice, s));
  ppm_bin_table_kernel<<<blocks_for(n_table, 256, 4096), 256, 0, s>>>(table_i, st.edges_dev.p, nb, energies, ntotal,
                                                                     n_occ, ppm_freq, naux);
  LAUNCH_CHECK();
  ppm_moments_kernel<<<dim3(slices, n_levels), kCmpMomentWarps * 32, 0, s>>>(
      M, ldn, slab, naux, energies, ppm_freq, ppm_fac, level_slab, table_i, st.edges_dev.p, nb,
      slices > 1 ? mom_part.p : st.mom.p);
  LAUNCH_CHECK();
  if (slices > 1) {
    sigma_ppm_grid_reduce<<<blocks_for(n_mom, 256, 2048), 256, 0, s>>>(st.mom.p, mom_part.p, n_mom, slices);
    LAUNCH_CHECK();
  }
  ppm_equivalent_poles_kernel<<<blocks_for(n_mom, 256, 2048), 256, 0, s>>>(
      reinterpret_cast<double2*>(st.eq.p), st.mom.p, st.edges_dev.p, nb, (long long)n_levels * nb);
  LAUNCH_CHECK();
  if (slices > 1) XTPB_CUDA(cudaStreamSynchronize(s));   // mom_part is freed on return
  st.valid = true;
}

// Compressed scan, part 2: n_items rows of n_omega uniformly spaced targets each (row i: slab level_slab[i], moment
// row level_mom[i] (nullptr: i), first target omega0[i]); near_host: [item][chunk][4] near / inner bin ranges.
void k_ppm_scan_evaluate(const double* M, long long ldn, long long slab, int naux, const double* energies,
                         const double* ppm_freq, const double* ppm_fac, const int* level_slab, const int* level_mom,
                         const double* omega0, double domega, int n_omega, int n_items, const int* near_host,
                         int n_chunks, const PpmScanState& st, double* values, double* direct_evaluations,
                         cudaStream_t s) {
  XTPB_REQUIRE(st.valid, "compressed scan evaluated without its prepared state");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int bx = (n_chunks + kCmpWarps - 1) / kCmpWarps;
  int splits = (int)((8LL * sms + (long long)bx * n_items - 1) / ((long long)bx * n_items));
  splits = std::max(1, std::min(splits, std::min(16, std::max(1, naux / 64))));
  const long long n = (long long)n_items * n_omega;
  const long long n_near = 4LL * n_items * n_chunks;
  DBuf near_buf((size_t)((n_near + 1) / 2)), partial, counter(1);
  counter.zero(s);
  if (splits > 1) partial.alloc((size_t)(n * splits));
  int* near_i = reinterpret_cast<int*>(near_buf.p);
  XTPB_CUDA(cudaMemcpyAsync(near_i, near_host, (size_t)n_near * sizeof(int), cudaMemcpyHostToDevice, s));
  const char* occ_env = std::getenv("XTPB_GRID_OCC");
  const int occ = occ_env ? std::atoi(occ_env) : kCmpMinBlocks;
  const char* walk_env = std::getenv("XTPB_GRID_WALK");
  const int walk = walk_env ? std::atoi(walk_env) : kCmpWalk;
  using KernelFn = decltype(&sigma_ppm_grid_compressed_kernel<5, 1>);
  auto pick = [&](auto walk_tag) -> KernelFn {
    constexpr int W = decltype(walk_tag)::value;
    return occ <= 3 ? sigma_ppm_grid_compressed_kernel<3, W>
         : occ == 4 ? sigma_ppm_grid_compressed_kernel<4, W>
         : occ == 5 ? sigma_ppm_grid_compressed_kernel<5, W>
                    : sigma_ppm_grid_compressed_kernel<6, W>;
  };
  KernelFn kernel = walk <= 0 ? pick(std::integral_constant<int, 0>{})
                  : walk == 1 ? pick(std::integral_constant<int, 1>{})
                              : pick(std::integral_constant<int, 2>{});
  // XTPB_GRID_KERNEL=cta: the CTA-cooperative near field (1c), possible whenever the union of the near bins of every
  // CTA fits its shared table.  Measured (profiles/r02_sigma_grid_walk.jsonl): 72.2 ms against 76.7 ms for the
  // warp-private kernel (1b) at synth-1000 size, no difference at C60 size (361 ms both) -- the staging it shares is a
  // minor part of the per-aux-function overhead -- so (1b) stays the default.
  const char* kern_env = std::getenv("XTPB_GRID_KERNEL");
  bool use_cta = kern_env && std::strcmp(kern_env, "cta") == 0;
  for (long long it = 0; it < n_items && use_cta; ++it)
    for (int c0 = 0; c0 < n_chunks && use_cta; c0 += kCmpWarps) {
      int lo = 0x7fffffff, hi = -1;
      for (int c = c0; c < std::min(n_chunks, c0 + kCmpWarps); ++c) {
        const int* nr = near_host + 4 * (it * n_chunks + c);
        if (nr[0] <= nr[1]) { lo = std::min(lo, nr[0]); hi = std::max(hi, nr[1]); }
      }
      if (hi >= lo && hi - lo + 2 > kCtaMaxTab) use_cta = false;
    }
  if (use_cta) {
    const int cocc = occ_env ? occ : kCtaMinBlocks;
    kernel = cocc <= 3 ? sigma_ppm_grid_cta_kernel<3> : cocc == 4 ? sigma_ppm_grid_cta_kernel<4>
                                                                  : sigma_ppm_grid_cta_kernel<5>;
  }
  for (int off = 0; off < n_items; off += 32768) {         // gridDim.y limit
    const int cnt = std::min(32768, n_items - off);
    kernel<<<dim3(bx, cnt, splits), kCmpWarps * 32, 0, s>>>(
        M, ldn, slab, naux, energies, ppm_freq, ppm_fac, level_slab + off, level_mom ? level_mom + off : nullptr,
        omega0 + off, domega, n_omega, reinterpret_cast<const int*>(st.table.p), st.edges_dev.p, st.nb,
        near_i + 4LL * off * n_chunks, n_chunks, st.mom.p, reinterpret_cast<const double2*>(st.eq.p),
        (splits > 1 ? partial.p : values) + (long long)off * n_omega, n,
        reinterpret_cast<unsigned long long*>(counter.p));
    LAUNCH_CHECK();
  }
  if (splits > 1) {
    sigma_ppm_grid_reduce<<<blocks_for(n, 256, 2048), 256, 0, s>>>(values, partial.p, n, splits);
    LAUNCH_CHECK();
  }
  unsigned long long near_poles = 0;
  XTPB_CUDA(cudaMemcpyAsync(&near_poles, counter.p, sizeof(near_poles), cudaMemcpyDeviceToHost, s));
  XTPB_CUDA(cudaStreamSynchronize(s));   // the scratch buffers are freed on return
  if (direct_evaluations) *direct_evaluations = (double)near_poles * kCmpChunk;
}

void k_sigma_ppm_grid_compressed(const double* M, long long ldn, long long slab, int ntotal, int naux, int n_occ,
                                 const double* energies, const double* ppm_freq, const double* ppm_fac,
                                 const int* level_slab, const double* omega0, double domega, int n_omega, int n_levels,
                                 const double* edges_host, int nb, const int* near_host, int n_chunks, double* values,
                                 double* direct_evaluations, PpmScanState& st, cudaStream_t s) {
  // work = pole evaluations of the equivalent direct sum (what (1) would do), so that rates stay comparable
  const int slot = prof_begin(PROF_SIGMA_GRID, (double)ntotal * naux * (double)n_omega * n_levels, s);
  k_ppm_scan_prepare(M, ldn, slab, ntotal, naux, n_occ, energies, ppm_freq, ppm_fac, level_slab, n_levels, edges_host,
                     nb, st, s);
  k_ppm_scan_evaluate(M, ldn, slab, naux, energies, ppm_freq, ppm_fac, level_slab, nullptr, omega0, domega, n_omega,
                      n_levels, near_host, n_chunks, st, values, direct_evaluations, s);
  prof_end(slot, s);
}
void k_sigma_ppm_pairs(const double* M, long long ldn, long long slab, int ntotal, int naux, int n_occ,
                       const double* energies, const double* ppm_freq, const double* ppm_fac, const int* pair_slab,
                       const double* pair_omega, int n_pairs, double* values, double* derivs, double* partial,
                       cudaStream_t s) {
  if (n_pairs == 0) return;
  const int slot = prof_begin(PROF_SIGMA_PAIRS, 8.0 * ntotal * (double)naux * n_pairs, s);   // work = slab bytes streamed
  for (int off = 0; off < n_pairs; off += 32768) {
    const int cnt = std::min(32768, n_pairs - off);
    sigma_ppm_pairs_kernel<<<dim3(kPairChunks, cnt), kPairThreads, 0, s>>>(
        M, ldn, slab, ntotal, naux, n_occ, energies, ppm_freq, ppm_fac, pair_slab + off, pair_omega + off,
        partial + (long long)off * kPairChunks * 2);
    LAUNCH_CHECK();
  }
  sigma_ppm_pairs_finalize<<<blocks_for(n_pairs, 128), 128, 0, s>>>(partial, n_pairs, values, derivs);
  LAUNCH_CHECK();
  prof_end(slot, s);
}
void k_sigma_ppm_weighted_slab(double* W, const double* M, long long ldn, long long slab, int ntotal, int p0, int pcnt,
                               int n_occ, const double* energies, const double* ppm_freq, const double* ppm_fac,
                               int slab0, int n_levels, const double* level_omega, cudaStream_t s) {
  const long long total = (long long)pcnt * ldn * n_levels;
  sigma_ppm_weighted_slab_kernel<<<blocks_for(total, 256, 16384), 256, 0, s>>>(
      W, M, ldn, slab, ntotal, p0, pcnt, n_occ, energies, ppm_freq, ppm_fac, slab0, n_levels, level_omega);
  LAUNCH_CHECK();
}
int sigma_ppm_pair_partial_doubles(int n_pairs) { return n_pairs * kPairChunks * 2; }

void k_bse_diagonal(double* diag, int vt, int ct, int naux, const double* Mvc, long long ldvc, long long slabvc,
                    const double* Mvv, long long ldvv, long long slabvv, const double* Mcc, long long ldcc,
                    long long slabcc, const double* Mcv, long long ldcv, long long slabcv, const double*,
                    const double* hqp_diag, int cqp, int cx, int cd, int cd2, cudaStream_t s) {
  DBuf dvv, dcc;
  if (cd) {
    dvv.alloc((size_t)vt * naux);
    dcc.alloc((size_t)ct * naux);
    bse_diag_helpers_kernel<<<blocks_for((long long)(vt + ct) * naux, 256, 4096), 256, 0, s>>>(
        dvv.p, dcc.p, vt, ct, naux, Mvv, ldvv, slabvv, Mcc, ldcc, slabcc);
    LAUNCH_CHECK();
  }
  bse_diagonal_kernel<<<vt, 128, 0, s>>>(diag, vt, ct, naux, Mvc, ldvc, slabvc, dvv.p, dcc.p, Mcv, ldcv, slabcv,
                                         hqp_diag, cqp, cx, cd, cd2);
  LAUNCH_CHECK();
  XTPB_CUDA(cudaStreamSynchronize(s));   // dvv/dcc are freed on return
}

void k_col_norms(const double* A, long long ld, long long rows, int cols, double* out, cudaStream_t s) {
  if (cols == 0) return;
  col_norms_kernel<<<cols, 256, 0, s>>>(A, ld, rows, out);
  LAUNCH_CHECK();
}
__global__ void scatter_fill_slots_kernel(double* __restrict__ M, long long ldn, long long slab,
                                          const double* __restrict__ stage, int mtotal, int ntotal, int naux, int world,
                                          int B, int round) {
  const long long slots = (long long)world * B;
  const long long total = (long long)mtotal * slots * ldn;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx % ldn);
    const long long t = (idx / ldn) % slots;
    const long long m = idx / (ldn * slots);
    if (n >= ntotal) continue;
    const int s = (int)(t / B), k = (int)(t % B);
    const long long a = (long long)naux * s / world, e = (long long)naux * (s + 1) / world;
    const long long P = a + (long long)round * B + k;
    if (P < e) M[m * slab + P * ldn + n] = stage[idx];
  }
}
void k_scatter_fill_slots(double* M, long long ldn, long long slab, const double* stage, int mtotal, int ntotal,
                          int naux, int world, int B, int round, cudaStream_t s) {
  const long long total = (long long)mtotal * world * B * ldn;
  scatter_fill_slots_kernel<<<blocks_for(total, 256, 16384), 256, 0, s>>>(M, ldn, slab, stage, mtotal, ntotal, naux,
                                                                        world, B, round);
  LAUNCH_CHECK();
}
void k_column_dots(double* out, const double* A, long long lda, const double* B, long long ldb, long long rows, int cols,
                   cudaStream_t s) {
  if (cols == 0) return;
  column_dots_kernel<<<cols, 256, 0, s>>>(A, lda, B, ldb, rows, out);
  LAUNCH_CHECK();
}
void k_residuals(double* res, long long ldr, const double* q, long long ldq, const double* lambda, long long rows,
                 int cols, cudaStream_t s) {
  residuals_kernel<<<blocks_for(rows * cols, 256, 4096), 256, 0, s>>>(res, ldr, q, ldq, lambda, rows, cols);
  LAUNCH_CHECK();
}
void k_davidson_correction(double* out, const double* r, const double* x, const double* D, double lambda, long long n,
                           int olsen, double* scratch2, cudaStream_t s) {
  // scratch2: n doubles (xd) followed by 2*nblocks partials
  const int nblocks = blocks_for(n, 256, 256);
  double* xd = scratch2;
  double* partial = scratch2 + n;
  correction_kernel<<<nblocks, 256, 0, s>>>(out, xd, r, x, D, lambda, n, olsen, partial);
  LAUNCH_CHECK();
  if (olsen) {
    olsen_finish_kernel<<<nblocks, 256, 0, s>>>(out, xd, partial, nblocks, n);
    LAUNCH_CHECK();
  }
}
void k_unpack_symmetric(double* full, long long ld, long long full_slice, const double* packed, long long pk_slice,
                        int n, int count, cudaStream_t s) {
  if (count == 0) return;
  const int t = (n + 31) / 32;
  const int slot = prof_begin(PROF_UNPACK, 12.0 * n * (double)n * count, s);
  for (int off = 0; off < count; off += 32768) {
    const int cnt = std::min(32768, count - off);
    unpack_symmetric_kernel<<<dim3(t * (t + 1) / 2, cnt), 256, 0, s>>>(full + (long long)off * full_slice, ld, full_slice,
                                                                      packed + (long long)off * pk_slice, pk_slice, n);
    LAUNCH_CHECK();
  }
  prof_end(slot, s);
}
void k_unit_vectors(double* V, long long ld, long long, const long long* idx, int cols, cudaStream_t s) {
  unit_vectors_kernel<<<blocks_for(cols, 128), 128, 0, s>>>(V, ld, idx, cols);
  LAUNCH_CHECK();
}

void k_gather_strided(double* dst, const double* src, long long first, long long stride, long long n, cudaStream_t s) {
  if (n == 0) return;
  gather_strided_kernel<<<blocks_for(n, 256, 1024), 256, 0, s>>>(dst, src, first, stride, n);
  LAUNCH_CHECK();
}
void k_cols_full_to_local(double* loc, long long ld_loc, const double* full, long long ld_full, long long rows,
                          long long ncols_loc, int rank, int world, cudaStream_t s) {
  cyclic_cols_kernel<true><<<blocks_for(rows * ncols_loc, 256, 8192), 256, 0, s>>>(loc, ld_loc, full, ld_full, rows,
                                                                                   ncols_loc, rank, world);
  LAUNCH_CHECK();
}
void k_cols_local_to_full(double* full, long long ld_full, const double* loc, long long ld_loc, long long rows,
                          long long ncols_loc, int rank, int world, cudaStream_t s) {
  cyclic_cols_kernel<false><<<blocks_for(rows * ncols_loc, 256, 8192), 256, 0, s>>>(full, ld_full, loc, ld_loc, rows,
                                                                                    ncols_loc, rank, world);
  LAUNCH_CHECK();
}
void k_window_from_gathered(double* dst, long long dst_ld, long long dst_slab, const double* G, long long ldl,
                            int mcnt, int naux, int P0, int pcnt, int n0, int ncnt, int world, cudaStream_t s) {
  const long long total = (long long)mcnt * pcnt * ncnt;
  if (total == 0) return;
  window_from_gathered_kernel<<<blocks_for(total, 256, 16384), 256, 0, s>>>(dst, dst_ld, dst_slab, G, ldl, mcnt, naux,
                                                                           P0, pcnt, n0, ncnt, world);
  LAUNCH_CHECK();
}

void k_bse_add_hqp(double* H, long long ld, int vt, int ct, int v2lo, int ns, const double* hqp, long long hs,
                   double cqp, cudaStream_t s) {
  for (int part = 0; part < 2; ++part) {
    const long long total = (long long)ns * ct * (part == 0 ? ct : vt);
    if (total == 0) continue;
    bse_add_hqp_kernel<<<blocks_for(total, 256, 8192), 256, 0, s>>>(H, ld, vt, ct, v2lo, ns, hqp, hs, cqp, part);
    LAUNCH_CHECK();
  }
}

void k_bse_pack_pairs(double* Ftri, long long ldt, const double* F, long long ldf, int vt, int naux, cudaStream_t s) {
  const long long npairs = (long long)vt * (vt + 1) / 2;
  bse_pack_pairs_kernel<<<blocks_for(npairs * naux, 256, 16384), 256, 0, s>>>(Ftri, ldt, F, ldf, vt, npairs, naux);
  LAUNCH_CHECK();
}
void k_bse_scatter_pairs(double* H, long long ld, int ct, const double* T, long long t0, long long tcnt, cudaStream_t s) {
  const int tiles = (ct + 31) / 32;
  for (long long off = 0; off < tcnt; off += 32768) {
    const long long cnt = std::min<long long>(32768, tcnt - off);
    bse_scatter_pairs_kernel<<<dim3(tiles * tiles, (unsigned)cnt), 256, 0, s>>>(H, ld, ct, T + off * (long long)ct * ct,
                                                                              t0 + off);
    LAUNCH_CHECK();
  }
}

}  // namespace xtpb
