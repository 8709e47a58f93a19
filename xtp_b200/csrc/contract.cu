// Launcher for the DMMA contraction engine: picks the tile configuration, split-K,
// alignment path; reduces split-K partials.
#include "contract.h"

#include <algorithm>
#include <cstdint>
#include <mutex>
#include <vector>

namespace xtpb {

std::atomic<long long> g_launch_count{0};
std::atomic<long long> g_tma_launch_count{0};
std::atomic<long long> g_tma5d_launch_count{0};

// ------------------------------------------------------------------ event-pair profiler
namespace {
struct ProfRec { cudaEvent_t e0, e1; int tag; double work; };
bool g_prof_on = false;
std::mutex g_prof_mu;                  // the Coulomb-metric prefetch thread may launch while the main thread does
std::vector<ProfRec> g_prof_pool;      // events are created once and reused after prof_reset()
size_t g_prof_used = 0;
constexpr size_t kProfCap = 1 << 17;
thread_local int t_prof_tag = PROF_OTHER;
}  // namespace

void prof_enable(bool on) { g_prof_on = on; }
bool prof_enabled() { return g_prof_on; }
void prof_reset() { g_prof_used = 0; }
int prof_begin(int tag, double work, cudaStream_t s) {
  if (!g_prof_on) return -1;
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if (g_prof_used >= kProfCap) return -1;
  if (g_prof_used == g_prof_pool.size()) {
    ProfRec r{};
    XTPB_CUDA(cudaEventCreate(&r.e0));
    XTPB_CUDA(cudaEventCreate(&r.e1));
    g_prof_pool.push_back(r);
  }
  ProfRec& r = g_prof_pool[g_prof_used];
  r.tag = tag < 0 ? t_prof_tag : tag;
  r.work = work;
  XTPB_CUDA(cudaEventRecord(r.e0, s));
  return (int)g_prof_used++;
}
void prof_end(int slot, cudaStream_t s) {
  if (slot < 0) return;
  cudaEvent_t e1;
  {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    e1 = g_prof_pool[(size_t)slot].e1;
  }
  XTPB_CUDA(cudaEventRecord(e1, s));
}
void prof_get(int tag, double* ms, double* work, long long* launches) {
  double tms = 0.0, tw = 0.0;
  long long n = 0;
  for (size_t i = 0; i < g_prof_used; ++i) {
    const ProfRec& r = g_prof_pool[i];
    if (r.tag != tag) continue;
    float t = 0.f;
    XTPB_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    tms += t;
    tw += r.work;
    ++n;
  }
  if (ms) *ms = tms;
  if (work) *work = tw;
  if (launches) *launches = n;
}
ProfScope::ProfScope(int tag) : prev(t_prof_tag) { t_prof_tag = tag; }
ProfScope::~ProfScope() { t_prof_tag = prev; }

namespace {

constexpr int kStages = 4;
constexpr int kMaxDevices = 64;
int g_num_sms[kMaxDevices] = {};

int current_device() {
  int dev = 0;
  XTPB_CUDA(cudaGetDevice(&dev));
  return dev;
}
// per-device state (one process may hold contexts on several devices): SM count, and which kernel instances have had
// their dynamic shared memory opt-in applied on which device
int num_sms() {
  const int dev = current_device();
  int& n = g_num_sms[dev % kMaxDevices];
  if (n == 0) XTPB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  return n;
}
template <class Kern>
void ensure_smem_optin(Kern kern, int bytes, unsigned long long& device_mask) {
  const int dev = current_device();
  const unsigned long long bit = 1ULL << (dev % kMaxDevices);
  if (device_mask & bit) return;
  XTPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  device_mask |= bit;
}

template <int BM, int BN, int WM, int WN, bool A_KC, bool B_KC, bool HAS_D, bool VEC>
void launch_cfg2(const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BM, BN, WM, WN, A_KC, B_KC, kStages>;
  auto kern = contract_kernel<Cfg, BM, BN, WM, WN, A_KC, B_KC, kStages, HAS_D, VEC>;
  static unsigned long long optin_mask = 0;
  ensure_smem_optin(kern, Cfg::SMEM_BYTES, optin_mask);
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
  dim3 grid(tiles_m * tiles_n, 1, p.n_batch * p.splits);
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(p);
  XTPB_CUDA(cudaGetLastError());
  ++g_launch_count;
}

// ---- TMA path: tensor maps are encoded per launch through the driver entry point (no link-time libcuda) ----
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}
bool tma_enabled() {
  static const bool on = [] { const char* e = getenv("XTPB_TMA"); return !(e && e[0] == '0'); }();
  return on && encode_tiled_fn() != nullptr;
}
// K-contig operand: dims (k, row, outer, batch), box (16, rows_box, 1, 1); row-contig: dims (row, k, outer, batch),
// box (16, 16, 1, 1).  Dimensions with a zero stride (broadcast) or extent 1 are dropped to extent 1 and the kernel
// multiplies their coordinate by 0.  Returns false when the operand cannot be described (falls back to cp.async).
bool make_tensor_map(CUtensorMap* map, const GemmOperand& o, bool kc, int rows, int rows_box, int K, int n_outer,
                     int n_batch, int* use_outer, int* use_batch) {
  *use_outer = (n_outer > 1 && o.s_outer != 0) ? 1 : 0;
  *use_batch = (n_batch > 1 && o.s_batch != 0) ? 1 : 0;
  if (reinterpret_cast<uintptr_t>(o.p) % 16) return false;
  const long long s_other = kc ? o.s_row : o.s_k;
  if (s_other <= 0 || s_other % 2 || (*use_outer && (o.s_outer <= 0 || o.s_outer % 2)) ||
      (*use_batch && (o.s_batch <= 0 || o.s_batch % 2)))
    return false;
  cuuint64_t dims[4] = {(cuuint64_t)(kc ? K : rows), (cuuint64_t)(kc ? rows : K),
                        (cuuint64_t)(*use_outer ? n_outer : 1), (cuuint64_t)(*use_batch ? n_batch : 1)};
  cuuint64_t strides[3] = {(cuuint64_t)s_other * 8, (cuuint64_t)(*use_outer ? o.s_outer * 8 : 16),
                           (cuuint64_t)(*use_batch ? o.s_batch * 8 : 16)};
  cuuint32_t box[4] = {16u, (cuuint32_t)(kc ? rows_box : 16), 1u, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  for (int i = 0; i < 3; ++i)
    if (strides[i] >= (1ULL << 40)) return false;
  const CUresult r = encode_tiled_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(o.p), dims, strides,
                                       box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}
// Row-contiguous operand as ONE box per tile: dims (row % 16, k, row / 16, outer, batch) with strides (8 B,) s_k, 128 B,
// ... and box (16, 16, rows_box / 16, 1, 1) land in shared memory as [row / 16][k][16 rows] -- exactly the mc tile
// layout the per-16-row boxes of the 4-D map produce, with one cp.async.bulk.tensor instead of rows_box / 16 of them
// (each costs the issuing warp ~200 cycles; with both operands row-contiguous the 16 issues per k-tile made warp 0 the
// straggler of its CTA: 29.6 TFLOP/s against 35.0 for two K-contiguous operands, profiles/r02_contract_diag.jsonl).
// The last row group of an operand whose row count is not a multiple of 16 reads up to 15 elements past the end of a
// k-row (the next k-row, or the slack every library allocation carries: core.cu device_alloc); those values only reach
// accumulators of rows >= M resp. columns >= N, which are never stored.  False: fall back to the 4-D map.
bool make_tensor_map_mc5(CUtensorMap* map, const GemmOperand& o, int rows, int rows_box, int K, int n_outer,
                         int n_batch, int* use_outer, int* use_batch) {
  static const bool on = [] { const char* e = getenv("XTPB_TMA5D"); return !(e && e[0] == '0'); }();
  if (!on || !o.library_owned) return false;
  *use_outer = (n_outer > 1 && o.s_outer != 0) ? 1 : 0;
  *use_batch = (n_batch > 1 && o.s_batch != 0) ? 1 : 0;
  if (reinterpret_cast<uintptr_t>(o.p) % 16) return false;
  if (o.s_k <= 0 || o.s_k % 2 || (*use_outer && (o.s_outer <= 0 || o.s_outer % 2)) ||
      (*use_batch && (o.s_batch <= 0 || o.s_batch % 2)))
    return false;
  cuuint64_t dims[5] = {16u, (cuuint64_t)K, (cuuint64_t)((rows + 15) / 16), (cuuint64_t)(*use_outer ? n_outer : 1),
                        (cuuint64_t)(*use_batch ? n_batch : 1)};
  cuuint64_t strides[4] = {(cuuint64_t)o.s_k * 8, 128u, (cuuint64_t)(*use_outer ? o.s_outer * 8 : 16),
                           (cuuint64_t)(*use_batch ? o.s_batch * 8 : 16)};
  cuuint32_t box[5] = {16u, 16u, (cuuint32_t)(rows_box / 16), 1u, 1u};
  cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
  for (int i = 0; i < 4; ++i)
    if (strides[i] >= (1ULL << 40)) return false;
  const CUresult r = encode_tiled_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, const_cast<double*>(o.p), dims, strides,
                                       box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BM, int BN, int WM, int WN, bool A_KC, bool B_KC, bool HAS_D>
bool launch_tma(const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BM, BN, WM, WN, A_KC, B_KC, kStages>;
  alignas(64) CUtensorMap mapA, mapB;
  TmaCoords tc{};
  tc.a_mc5 = !A_KC && make_tensor_map_mc5(&mapA, p.A, p.M, BM, p.K, p.n_outer, p.n_batch, &tc.a_outer, &tc.a_batch);
  tc.b_mc5 = !B_KC && make_tensor_map_mc5(&mapB, p.B, p.N, BN, p.K, p.n_outer, p.n_batch, &tc.b_outer, &tc.b_batch);
  if (!tc.a_mc5 && !make_tensor_map(&mapA, p.A, A_KC, p.M, BM, p.K, p.n_outer, p.n_batch, &tc.a_outer, &tc.a_batch))
    return false;
  if (!tc.b_mc5 && !make_tensor_map(&mapB, p.B, B_KC, p.N, BN, p.K, p.n_outer, p.n_batch, &tc.b_outer, &tc.b_batch))
    return false;
  if (tc.a_mc5 || tc.b_mc5) ++g_tma5d_launch_count;
  auto kern = contract_tma_kernel<Cfg, BM, BN, WM, WN, A_KC, B_KC, kStages, HAS_D>;
  constexpr int smem_bytes = kStages * (Cfg::A_BYTES + Cfg::B_BYTES + 1024) + 2 * kStages * 8;
  static unsigned long long optin_mask = 0;
  ensure_smem_optin(kern, smem_bytes, optin_mask);
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
  dim3 grid(tiles_m * tiles_n, 1, p.n_batch * p.splits);
  kern<<<grid, Cfg::THREADS, smem_bytes, stream>>>(p, tc, mapA, mapB);
  XTPB_CUDA(cudaGetLastError());
  ++g_launch_count;
  return true;
}

// TMA instance when both operands are 16-byte aligned with even strides (and the problem is big enough to amortise
// the two descriptor encodes); 16-byte cp.async instance otherwise; 8-byte cp.async instance for odd alignments.
template <int BM, int BN, int WM, int WN, bool A_KC, bool B_KC, bool HAS_D>
void launch_cfg(const GemmParams& p, cudaStream_t stream) {
  if (p.a_vec && p.b_vec) {
    const double work = (double)p.M * p.N * (double)p.K * p.n_outer * p.n_batch;
    if (p.use_tma && (p.use_tma == 2 || work >= 1.0e8) && tma_enabled() &&
        launch_tma<BM, BN, WM, WN, A_KC, B_KC, HAS_D>(p, stream)) {
      ++g_tma_launch_count;
      return;
    }
    launch_cfg2<BM, BN, WM, WN, A_KC, B_KC, HAS_D, true>(p, stream);
  } else {
    launch_cfg2<BM, BN, WM, WN, A_KC, B_KC, HAS_D, false>(p, stream);
  }
}

template <bool A_KC, bool B_KC, bool HAS_D>
void launch_tile(const GemmParams& p, int cfg, cudaStream_t stream) {
  switch (cfg) {
    case 0: launch_cfg<128, 128, 64, 32, A_KC, B_KC, HAS_D>(p, stream); break;
    case 1: launch_cfg<128, 64, 64, 32, A_KC, B_KC, HAS_D>(p, stream); break;
    default: launch_cfg<128, 32, 32, 32, A_KC, B_KC, HAS_D>(p, stream); break;
  }
}

template <bool A_KC, bool B_KC>
void launch_d(const GemmParams& p, int cfg, cudaStream_t stream) {
  if (p.d) launch_tile<A_KC, B_KC, true>(p, cfg, stream);
  else launch_tile<A_KC, B_KC, false>(p, cfg, stream);
}

__global__ void splitk_reduce_kernel(const double* __restrict__ ws, double* __restrict__ C, int M, int N, int splits,
                                     long long c_sm, long long c_sn, long long c_batch, int c_n_inner,
                                     long long c_sn_outer, int c_m_inner, long long c_sm_outer, double alpha,
                                     double beta, int lower) {
  const long long mn = (long long)M * N;
  const int batch = blockIdx.y;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < mn;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % M), c = (int)(idx / M);
    if (lower && c > r) continue;
    const double* w = ws + (long long)batch * splits * mn + idx;
    double s = 0.0;
    for (int k = 0; k < splits; ++k) s += w[(long long)k * mn];
    const long long coff = c_n_inner > 0 ? (long long)(c / c_n_inner) * c_sn_outer + (long long)(c % c_n_inner) * c_sn
                                         : (long long)c * c_sn;
    const long long roff = c_m_inner > 0 ? (long long)(r / c_m_inner) * c_sm_outer + (long long)(r % c_m_inner) * c_sm
                                         : (long long)r * c_sm;
    double* dst = C + (long long)batch * c_batch + roff + coff;
    double v = alpha * s;
    if (beta != 0.0) v += beta * (*dst);
    *dst = v;
  }
}

__global__ void symmetrize_kernel(double* C, int n, long long ld, double diag_add) {
  __shared__ double tile[32][33];
  const int bi = blockIdx.x, bj = blockIdx.y;   // tile (bi, bj) of the lower triangle, bi >= bj
  if (bj > bi) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  // read lower tile: rows bi*32.., cols bj*32..
  for (int cc = ty; cc < 32; cc += blockDim.y) {
    const int r = bi * 32 + tx, c = bj * 32 + cc;
    tile[cc][tx] = (r < n && c < n) ? C[r + c * ld] : 0.0;
  }
  __syncthreads();
  // write transposed into the upper tile: rows bj*32.., cols bi*32..
  for (int cc = ty; cc < 32; cc += blockDim.y) {
    const int r = bj * 32 + tx, c = bi * 32 + cc;   // element (r, c) = lower(c, r) = tile[r - bj*32][c - bi*32]
    if (r < n && c < n) {
      if (c > r) C[r + c * ld] = tile[tx][cc];
      else if (c == r && diag_add != 0.0) C[r + c * ld] = tile[tx][cc] + diag_add;
    }
  }
}

bool aligned16(const GemmOperand& o, bool kc, int n_outer, int n_batch) {
  if (reinterpret_cast<uintptr_t>(o.p) % 16) return false;
  const long long other = kc ? o.s_row : o.s_k;
  if (other % 2) return false;
  if (n_outer > 1 && o.s_outer % 2) return false;
  if (n_batch > 1 && o.s_batch % 2) return false;
  return true;
}

}  // namespace

// Launch plan of a contraction (host arithmetic only, no device): tile configuration `cfg` (0: 128x128, 1: 128x64,
// 2: 128x32; < 0 on entry = choose) and split-K factor `splits` (<= 0 on entry = choose) for `sms` SMs.
void contract_plan(int M, int N, int K, int n_outer, int n_batch, int lower, int sms, int& cfg, int& splits) {
  if (cfg < 0) {
    if (N <= 32) cfg = 2;
    else if (N <= 64) cfg = 1;
    else {
      const long long pad128 = round_up(N, 128), pad64 = round_up(N, 64);
      cfg = (pad64 < pad128) ? 1 : 0;
    }
    if (lower) cfg = 0;
  }
  const int bn = cfg == 0 ? 128 : (cfg == 1 ? 64 : 32);
  const long long tiles = (long long)((M + 127) / 128) * ((N + bn - 1) / bn) * n_batch;
  const long long nkt = (long long)((K + BK - 1) / BK) * n_outer;
  if (splits > 0) return;
  splits = 1;
  // fewer than two waves of tiles: split the contraction index so that ~3 waves of CTAs are in flight (tail
  // balance for the tensor-bound shapes, bytes in flight for the HBM-bound tall-skinny ones)
  if (tiles < 2LL * sms && nkt >= 16) {
    splits = (int)std::min<long long>({(3LL * sms + tiles - 1) / tiles, nkt / 8, 64LL});
    if (splits < 1) splits = 1;
  } else if (nkt >= 512) {
    // Long contractions with a handful of waves (the eps(w) SYRK at C60 size: 946 lower-triangle tiles = 6.4 waves of
    // 40 ms tiles, of which the seventh runs 0.4 full): splitting the contraction index s ways turns the tail into
    // ceil(s * tiles / slots) / s waves.  Taken when the saved tile time clearly exceeds the extra pass over the
    // partial results (and the workspace stays below 2 GiB); deterministic like every split-K launch.
    long long real_tiles = tiles;
    if (lower) {          // tiles on or below the diagonal (BM = BN = 128 for lower-triangular outputs)
      const long long tm = (M + 127) / 128, tn = (N + 127) / 128;
      real_tiles = 0;
      for (long long j = 0; j < tn; ++j) real_tiles += std::max<long long>(0, tm - j);
      real_tiles *= n_batch;
    }
    const long long slots = (long long)sms * (cfg == 0 ? 1 : 2);
    auto waves = [&](int s_) { return double((real_tiles * s_ + slots - 1) / slots) / s_; };
    const double tile_us = 2.1 * double(nkt) * (bn / 128.0);             // one CTA, whole contraction index
    const double mn_bytes = 8.0 * double(M) * double(N) * n_batch * (lower ? 0.5 : 1.0);
    int best = 1;
    double best_gain = 0.0;
    for (int s_ = 2; s_ <= 8; ++s_) {
      if (nkt / s_ < 256 || mn_bytes * (lower ? 2.0 : 1.0) * s_ > 2147483648.0) break;
      const double saved_us = (waves(1) - waves(s_)) * tile_us;
      const double reduce_us = (s_ + 1) * mn_bytes / 4.0e6 + 10.0;       // ~4 TB/s over partials + output
      const double gain = saved_us - 3.0 * reduce_us;
      if (gain > best_gain + 1e-9) { best_gain = gain; best = s_; }
    }
    static const bool tail_split = [] { const char* e = getenv("XTPB_TAIL_SPLIT"); return !(e && e[0] == '0'); }();
    if (tail_split) splits = best;
  }
}

int contract(GemmParams p, Workspace& ws, cudaStream_t stream, int force_cfg, int force_splits) {
  XTPB_REQUIRE(p.M > 0 && p.N > 0 && p.K >= 0 && p.n_outer >= 1 && p.n_batch >= 1, "bad contraction sizes");
  XTPB_REQUIRE((p.A.s_row == 1) != (p.A.s_k == 1) || (p.A.s_row == 1 && p.A.s_k == 1 && (p.M == 1 || p.K == 1)) ||
                   (p.A.s_row == 1 && p.A.s_k == 1),
               "operand A needs a unit stride");
  const bool a_kc = p.A.s_k == 1;
  const bool b_kc = p.B.s_k == 1;
  XTPB_REQUIRE(a_kc || p.A.s_row == 1, "operand A needs a unit stride");
  XTPB_REQUIRE(b_kc || p.B.s_row == 1, "operand B needs a unit stride");
  p.use_tma = 1;
  if (force_cfg >= 8) {        // 8..10: tile 0..2 on the TMA instance whatever the problem size (tests)
    p.use_tma = 2;
    force_cfg -= 8;
  } else if (force_cfg >= 4) { // 4..6: tile 0..2 on the cp.async instance, 7: automatic tile on cp.async (tests, A/B runs)
    p.use_tma = 0;
    force_cfg = force_cfg == 7 ? -1 : force_cfg - 4;
  }
  p.a_vec = aligned16(p.A, a_kc, p.n_outer, p.n_batch) ? 1 : 0;
  p.b_vec = aligned16(p.B, b_kc, p.n_outer, p.n_batch) ? 1 : 0;

  int cfg = force_cfg, splits = force_splits;
  contract_plan(p.M, p.N, p.K, p.n_outer, p.n_batch, p.lower, num_sms(), cfg, splits);
  XTPB_REQUIRE((long long)p.n_batch * splits <= 65535, "batch*splits exceeds gridDim.z");
  p.splits = splits;
  p.ws = nullptr;
  if (splits > 1) p.ws = ws.get((size_t)p.n_batch * splits * p.M * p.N);

  // algorithmic flops: 2 M N K_total, halved for the lower-triangular (SYRK-style) outputs
  const double flops = (p.lower ? 1.0 : 2.0) * (double)p.M * p.N * (double)p.K * p.n_outer * p.n_batch;
  const int prof_slot = prof_begin(-1, flops, stream);
  if (a_kc && b_kc) launch_d<true, true>(p, cfg, stream);
  else if (a_kc && !b_kc) launch_d<true, false>(p, cfg, stream);
  else if (!a_kc && b_kc) launch_d<false, true>(p, cfg, stream);
  else launch_d<false, false>(p, cfg, stream);

  int launches = 1;
  if (splits > 1) {
    const long long mn = (long long)p.M * p.N;
    dim3 grid((unsigned)std::min<long long>((mn + 255) / 256, 4096), p.n_batch);
    splitk_reduce_kernel<<<grid, 256, 0, stream>>>(p.ws, p.C, p.M, p.N, splits, p.c_sm, p.c_sn, p.c_batch, p.c_n_inner,
                                                   p.c_sn_outer, p.c_m_inner, p.c_sm_outer, p.alpha, p.beta, p.lower);
    XTPB_CUDA(cudaGetLastError());
    ++g_launch_count;
    ++launches;
  }
  prof_end(prof_slot, stream);
  return launches;
}

void symmetrize_from_lower(double* C, int n, long long ld, double diag_add, cudaStream_t stream) {
  const int t = (n + 31) / 32;
  symmetrize_kernel<<<dim3(t, t), dim3(32, 8), 0, stream>>>(C, n, ld, diag_add);
  XTPB_CUDA(cudaGetLastError());
  ++g_launch_count;
}

}  // namespace xtpb
