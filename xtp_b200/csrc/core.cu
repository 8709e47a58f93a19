// Context, error reporting, dense solver wrappers (cuSOLVER eigh / inverse; these are the
// "serial-ish library calls" of SURVEY.md section 7 hard part 6 and are timed apart from the contractions).
#include <chrono>
#include <mutex>
#include <map>
#include <utility>
#include <vector>

#include "internal.h"

namespace xtpb {

namespace {
thread_local std::string g_last_error;
}
void set_last_error(const std::string& msg) { g_last_error = msg; }

// ---------------------------------------------------------------- device memory (see common.h)
namespace {
std::mutex g_alloc_mu;
std::multimap<std::pair<int, size_t>, void*> g_block_cache;   // (device, exact size) -> released block
std::pair<int, size_t> cache_key(size_t bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  return {dev, bytes};
}
double g_cached_bytes = 0.0, g_alloc_seconds = 0.0, g_free_wait_seconds = 0.0;
long long g_alloc_calls = 0, g_cache_hits = 0;
// on unless XTPB_ALLOC_CACHE=0; at most XTPB_ALLOC_CACHE_MAX_GB (default 64) are kept
bool alloc_cache_on() {
  static const bool on = [] { const char* e = getenv("XTPB_ALLOC_CACHE"); return !(e && e[0] == '0'); }();
  return on;
}
double alloc_cache_cap_bytes() {
  static const double cap = [] {
    const char* e = getenv("XTPB_ALLOC_CACHE_MAX_GB");
    return (e ? atof(e) : 64.0) * 1e9;
  }();
  return cap;
}
double seconds_since(std::chrono::steady_clock::time_point t0) {
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
void flush_block_cache_locked() {
  for (auto& kv : g_block_cache) cudaFree(kv.second);
  g_block_cache.clear();
  g_cached_bytes = 0.0;
}
}  // namespace

constexpr size_t kAllocSlack = 256;

void* device_alloc(size_t bytes) {
  const auto t0 = std::chrono::steady_clock::now();
  void* p = nullptr;
  std::lock_guard<std::mutex> lock(g_alloc_mu);
  if (alloc_cache_on()) {
    auto it = g_block_cache.find(cache_key(bytes));
    if (it != g_block_cache.end()) {
      p = it->second;
      g_block_cache.erase(it);
      g_cached_bytes -= (double)bytes;
      ++g_cache_hits;
    }
  }
  if (!p) {
    // kAllocSlack bytes behind every block: single-box TMA loads of row-contiguous operands may read up to 15 doubles
    // past the last element of an operand (contract.cu: make_tensor_map_mc5)
    cudaError_t err = cudaMalloc(&p, bytes + kAllocSlack);
    if (err == cudaErrorMemoryAllocation && !g_block_cache.empty()) {
      cudaGetLastError();                       // clear the sticky error, give the cached blocks back, try again
      cudaDeviceSynchronize();
      flush_block_cache_locked();
      err = cudaMalloc(&p, bytes + kAllocSlack);
    }
    if (err != cudaSuccess)
      throw Error(std::string("CUDA error ") + cudaGetErrorString(err) + " in cudaMalloc of " + std::to_string(bytes) +
                  " bytes");
  }
  g_alloc_seconds += seconds_since(t0);
  ++g_alloc_calls;
  return p;
}

void device_free(void* p, size_t bytes) {
  if (!p) return;
  const auto t0 = std::chrono::steady_clock::now();
  std::lock_guard<std::mutex> lock(g_alloc_mu);
  if (alloc_cache_on() && g_cached_bytes + (double)bytes <= alloc_cache_cap_bytes()) {
    // what cudaFree implies: nothing in flight may still use the block.  The wait is for outstanding GPU work (e.g. an
    // eigensolver running on the helper stream), not allocator time: it is accounted separately.
    cudaDeviceSynchronize();
    g_free_wait_seconds += seconds_since(t0);
    const auto t1 = std::chrono::steady_clock::now();
    g_block_cache.emplace(cache_key(bytes), p);
    g_cached_bytes += (double)bytes;
    g_alloc_seconds += seconds_since(t1);
    return;
  }
  cudaFree(p);
  g_alloc_seconds += seconds_since(t0);
}

double device_free_wait_seconds() {
  std::lock_guard<std::mutex> lock(g_alloc_mu);
  return g_free_wait_seconds;
}
void device_alloc_stats(double* seconds, long long* calls, long long* cache_hits, double* cached_bytes, bool reset) {
  std::lock_guard<std::mutex> lock(g_alloc_mu);
  if (seconds) *seconds = g_alloc_seconds;
  if (calls) *calls = g_alloc_calls;
  if (cache_hits) *cache_hits = g_cache_hits;
  if (cached_bytes) *cached_bytes = g_cached_bytes;
  if (reset) {
    g_free_wait_seconds = 0.0;
    g_alloc_seconds = 0.0;
    g_alloc_calls = 0;
    g_cache_hits = 0;
  }
}
const char* last_error_cstr() { return g_last_error.c_str(); }

#define XTPB_SOLVER(expr)                                                                       \
  do {                                                                                          \
    cusolverStatus_t st__ = (expr);                                                             \
    if (st__ != CUSOLVER_STATUS_SUCCESS)                                                        \
      throw ::xtpb::Error(std::string("cuSOLVER status ") + std::to_string((int)st__) + " in " #expr); \
  } while (0)

Context::Context(int dev) : device(dev) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    throw Error("xtpb: no CUDA device available (libxtpb200 has no CPU fallback)");
  XTPB_REQUIRE(dev >= 0 && dev < count, "device index out of range");
  XTPB_CUDA(cudaSetDevice(dev));
  cudaDeviceProp prop;
  XTPB_CUDA(cudaGetDeviceProperties(&prop, dev));
  XTPB_REQUIRE(prop.major >= 10, "libxtpb200 is built for sm_100a (Blackwell) only");
  XTPB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  XTPB_SOLVER(cusolverDnCreate(&solver));
  XTPB_SOLVER(cusolverDnSetStream(solver, stream));
  XTPB_CUDA(cudaMalloc(&dev_info, sizeof(int)));
  XTPB_CUDA(cudaEventCreate(&ev0));
  XTPB_CUDA(cudaEventCreate(&ev1));
}

Context::~Context() {
  comm_destroy();
  if (async_eigh.th.joinable()) async_eigh.th.join();
  if (side_ready) cudaEventDestroy(side_ready);
  if (side_solver) cusolverDnDestroy(side_solver);
  if (side_info) cudaFree(side_info);
  if (side_stream) cudaStreamDestroy(side_stream);
  side_work.release();
  solver_work.release();
  scratch_a.release();
  scratch_b.release();
  ws.buf.release();
  {   // blocks cached for this device go back to the driver with the context
    std::lock_guard<std::mutex> lock(g_alloc_mu);
    cudaDeviceSynchronize();
    for (auto it = g_block_cache.begin(); it != g_block_cache.end();) {
      if (it->first.first == device) {
        cudaFree(it->second);
        g_cached_bytes -= (double)it->first.second;
        it = g_block_cache.erase(it);
      } else {
        ++it;
      }
    }
  }
  if (solver) cusolverDnDestroy(solver);
  if (dev_info) cudaFree(dev_info);
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  if (stream) cudaStreamDestroy(stream);
}

void Context::solver_begin() {
  XTPB_CUDA(cudaEventRecord(ev0, stream));
  solver_prof_slot = prof_begin(PROF_SOLVER, 0.0, stream);
}
void Context::solver_end() {
  prof_end(solver_prof_slot, stream);
  XTPB_CUDA(cudaEventRecord(ev1, stream));
  XTPB_CUDA(cudaEventSynchronize(ev1));
  float ms = 0;
  XTPB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
  solver_seconds += ms * 1e-3;
  int info = 0;
  XTPB_CUDA(cudaMemcpy(&info, dev_info, sizeof(int), cudaMemcpyDeviceToHost));
  XTPB_REQUIRE(info == 0, "cuSOLVER reported a non-zero devInfo (" + std::to_string(info) + ")");
}

void Context::eigh(int n, double* A, long long lda, double* w) {
  int lwork = 0;
  XTPB_SOLVER(cusolverDnDsyevd_bufferSize(solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A, (int)lda, w,
                                          &lwork));
  solver_work.ensure((size_t)lwork);
  solver_begin();
  XTPB_SOLVER(cusolverDnDsyevd(solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A, (int)lda, w,
                               solver_work.p, lwork, dev_info));
  solver_end();
}

void Context::spd_inverse(int n, double* A, long long lda) {
  int lwork = 0, lwork2 = 0;
  XTPB_SOLVER(cusolverDnDpotrf_bufferSize(solver, CUBLAS_FILL_MODE_LOWER, n, A, (int)lda, &lwork));
  XTPB_SOLVER(cusolverDnDpotri_bufferSize(solver, CUBLAS_FILL_MODE_LOWER, n, A, (int)lda, &lwork2));
  solver_work.ensure((size_t)std::max(lwork, lwork2));
  solver_begin();
  XTPB_SOLVER(cusolverDnDpotrf(solver, CUBLAS_FILL_MODE_LOWER, n, A, (int)lda, solver_work.p, lwork, dev_info));
  solver_end();
  solver_begin();
  XTPB_SOLVER(cusolverDnDpotri(solver, CUBLAS_FILL_MODE_LOWER, n, A, (int)lda, solver_work.p, lwork2, dev_info));
  solver_end();
  symmetrize_from_lower(A, n, lda, 0.0, stream);
}

bool Context::cholesky(int n, double* A, long long lda, bool upper) {
  const cublasFillMode_t uplo = upper ? CUBLAS_FILL_MODE_UPPER : CUBLAS_FILL_MODE_LOWER;
  int lwork = 0;
  XTPB_SOLVER(cusolverDnDpotrf_bufferSize(solver, uplo, n, A, (int)lda, &lwork));
  solver_work.ensure((size_t)lwork);
  XTPB_CUDA(cudaEventRecord(ev0, stream));
  solver_prof_slot = prof_begin(PROF_SOLVER, 0.0, stream);
  XTPB_SOLVER(cusolverDnDpotrf(solver, uplo, n, A, (int)lda, solver_work.p, lwork, dev_info));
  prof_end(solver_prof_slot, stream);
  XTPB_CUDA(cudaEventRecord(ev1, stream));
  XTPB_CUDA(cudaEventSynchronize(ev1));
  float ms = 0;
  XTPB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
  solver_seconds += ms * 1e-3;
  int info = 0;
  XTPB_CUDA(cudaMemcpy(&info, dev_info, sizeof(int), cudaMemcpyDeviceToHost));
  XTPB_REQUIRE(info >= 0, "cuSOLVER potrf: illegal argument");
  return info == 0;
}

void Context::tri_inverse(int n, double* L, long long lda, bool upper) {
  const cublasFillMode_t uplo = upper ? CUBLAS_FILL_MODE_UPPER : CUBLAS_FILL_MODE_LOWER;
  size_t dev_bytes = 0, host_bytes = 0;
  XTPB_SOLVER(cusolverDnXtrtri_bufferSize(solver, uplo, CUBLAS_DIAG_NON_UNIT, n, CUDA_R_64F, L, lda,
                                          &dev_bytes, &host_bytes));
  solver_work.ensure((dev_bytes + 7) / 8 + 1);
  std::vector<char> host_work(host_bytes + 1);
  solver_begin();
  XTPB_SOLVER(cusolverDnXtrtri(solver, uplo, CUBLAS_DIAG_NON_UNIT, n, CUDA_R_64F, L, lda,
                               solver_work.p, dev_bytes, host_work.data(), host_bytes, dev_info));
  solver_end();
}

void Context::general_inverse(int n, double* A, long long lda, double* Ainv, long long ldi) {
  int lwork = 0;
  XTPB_SOLVER(cusolverDnDgetrf_bufferSize(solver, n, n, A, (int)lda, &lwork));
  solver_work.ensure((size_t)lwork + (size_t)(n + 1) / 2 + 1);
  int* ipiv = reinterpret_cast<int*>(solver_work.p + lwork);
  k_set_identity(Ainv, n, ldi, stream);
  solver_begin();
  XTPB_SOLVER(cusolverDnDgetrf(solver, n, n, A, (int)lda, solver_work.p, ipiv, dev_info));
  solver_end();
  solver_begin();
  XTPB_SOLVER(cusolverDnDgetrs(solver, CUBLAS_OP_N, n, n, A, (int)lda, ipiv, Ainv, (int)ldi, dev_info));
  solver_end();
}

}  // namespace xtpb

namespace xtpb {
// x <- A^-1 x for one right-hand side, A symmetric (destroyed).  The CDA residues need v^T eps^-1 v for ONE vector per
// matrix: a factorisation + one pair of triangular solves instead of a full inverse (8/3 n^3).  eps(w) on the real
// axis is positive definite below the first transition energy, so Cholesky (n^3/3, and cuSOLVER's potrf runs several
// times faster than its getrf) is tried first on a copy; an indefinite matrix (devInfo > 0) falls back to LU.
void Context::lu_solve_vector(int n, double* A, long long lda, double* x) {
  int lwork = 0, lwork_c = 0;
  XTPB_SOLVER(cusolverDnDgetrf_bufferSize(solver, n, n, A, (int)lda, &lwork));
  XTPB_SOLVER(cusolverDnDpotrf_bufferSize(solver, CUBLAS_FILL_MODE_LOWER, n, A, (int)lda, &lwork_c));
  lwork = std::max(lwork, lwork_c);
  solver_work.ensure((size_t)lwork + (size_t)(n + 1) / 2 + 1);
  scratch_a.ensure((size_t)n * n);
  int* ipiv = reinterpret_cast<int*>(solver_work.p + lwork);
  XTPB_CUDA(cudaMemcpy2DAsync(scratch_a.p, (size_t)n * 8, A, (size_t)lda * 8, (size_t)n * 8, n, cudaMemcpyDeviceToDevice,
                              stream));
  XTPB_CUDA(cudaEventRecord(ev0, stream));
  solver_prof_slot = prof_begin(PROF_SOLVER, 0.0, stream);
  XTPB_SOLVER(cusolverDnDpotrf(solver, CUBLAS_FILL_MODE_LOWER, n, scratch_a.p, n, solver_work.p, lwork, dev_info));
  int info = 0;
  XTPB_CUDA(cudaMemcpyAsync(&info, dev_info, sizeof(int), cudaMemcpyDeviceToHost, stream));
  XTPB_CUDA(cudaStreamSynchronize(stream));
  if (info == 0) {
    XTPB_SOLVER(cusolverDnDpotrs(solver, CUBLAS_FILL_MODE_LOWER, n, 1, scratch_a.p, n, x, n, dev_info));
  } else {
    XTPB_REQUIRE(info > 0, "cuSOLVER potrf: illegal argument");
    XTPB_SOLVER(cusolverDnDgetrf(solver, n, n, A, (int)lda, solver_work.p, ipiv, dev_info));
    XTPB_SOLVER(cusolverDnDgetrs(solver, CUBLAS_OP_N, n, 1, A, (int)lda, ipiv, x, n, dev_info));
  }
  solver_end();
}
}  // namespace xtpb

namespace xtpb {
void Context::side_init() {
  if (side_stream) return;
  int lo = 0, hi = 0;
  XTPB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  XTPB_CUDA(cudaStreamCreateWithPriority(&side_stream, cudaStreamNonBlocking, hi));
  XTPB_SOLVER(cusolverDnCreate(&side_solver));
  XTPB_SOLVER(cusolverDnSetStream(side_solver, side_stream));
  XTPB_CUDA(cudaMalloc(&side_info, sizeof(int)));
  XTPB_CUDA(cudaEventCreateWithFlags(&side_ready, cudaEventDisableTiming));
}

void Context::eigh_async_begin(int n, double* A, long long lda, double* w, double* lam_host) {
  side_init();
  if (async_eigh.th.joinable()) async_eigh.th.join();
  async_eigh.err = nullptr;
  async_eigh.active = true;
  int lwork = 0;
  XTPB_SOLVER(cusolverDnDsyevd_bufferSize(side_solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A, (int)lda,
                                          w, &lwork));
  side_work.ensure((size_t)lwork);
  XTPB_CUDA(cudaEventRecord(side_ready, stream));            // A is complete once the main stream reaches this point
  XTPB_CUDA(cudaStreamWaitEvent(side_stream, side_ready, 0));
  async_eigh.th = std::thread([this, n, A, lda, w, lam_host, lwork] {
    try {
      XTPB_CUDA(cudaSetDevice(device));
      XTPB_SOLVER(cusolverDnDsyevd(side_solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A, (int)lda, w,
                                   side_work.p, lwork, side_info));
      int info = 0;
      XTPB_CUDA(cudaMemcpyAsync(lam_host, w, (size_t)n * 8, cudaMemcpyDeviceToHost, side_stream));
      XTPB_CUDA(cudaMemcpyAsync(&info, side_info, sizeof(int), cudaMemcpyDeviceToHost, side_stream));
      XTPB_CUDA(cudaStreamSynchronize(side_stream));
      XTPB_REQUIRE(info == 0, "cuSOLVER devInfo " + std::to_string(info) + " in the overlapped eigensolver");
    } catch (...) {
      async_eigh.err = std::current_exception();
    }
  });
}

void Context::eigh_async_join() {
  if (!async_eigh.active) return;
  if (async_eigh.th.joinable()) async_eigh.th.join();
  async_eigh.active = false;
  if (async_eigh.err) std::rethrow_exception(async_eigh.err);
}
}  // namespace xtpb
