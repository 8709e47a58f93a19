// Shared host-side plumbing for libxtpb200: error reporting across the C ABI,
// RAII device buffers.  No exceptions cross the C boundary: every extern "C"
// entry point wraps its body in XTPB_API_BEGIN/END and returns an int status.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

namespace xtpb {

void set_last_error(const std::string& msg);

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define XTPB_CUDA(expr)                                                                          \
  do {                                                                                           \
    cudaError_t err__ = (expr);                                                                  \
    if (err__ != cudaSuccess) {                                                                  \
      throw ::xtpb::Error(std::string("CUDA error ") + cudaGetErrorString(err__) + " at " +      \
                          __FILE__ + ":" + std::to_string(__LINE__) + " in " #expr);             \
    }                                                                                            \
  } while (0)

#define XTPB_REQUIRE(cond, msg)                                                                  \
  do {                                                                                           \
    if (!(cond)) throw ::xtpb::Error(std::string("xtpb: ") + (msg) + " [" #cond "]");            \
  } while (0)

#define XTPB_API_BEGIN try {
#define XTPB_API_END                                   \
  return 0;                                            \
  }                                                    \
  catch (const std::exception& e) {                    \
    ::xtpb::set_last_error(e.what());                  \
    return 1;                                          \
  }                                                    \
  catch (...) {                                        \
    ::xtpb::set_last_error("unknown C++ exception");   \
    return 2;                                          \
  }

// Device memory for DBuf (core.cu).  Host seconds spent inside cudaMalloc / cudaFree are accumulated (bench.py reports
// them per step: they are pure host/driver time between kernels).  Unless XTPB_ALLOC_CACHE=0, released blocks are kept
// in an exact-size cache (capped by XTPB_ALLOC_CACHE_MAX_GB, default 64) instead of going back to the driver: a step
// allocates the same sizes again and again
// (scratch for rotations, epsilon, the BSE operands), so after the first step almost every allocation is a cache hit.
// release() then performs the device-wide synchronisation cudaFree would have implied, so that stream-ordering
// assumptions of the callers are unchanged; the cache is flushed when cudaMalloc runs out of memory.
void* device_alloc(size_t bytes);
void device_free(void* p, size_t bytes);
void device_alloc_stats(double* seconds, long long* calls, long long* cache_hits, double* cached_bytes, bool reset);
// host seconds device_free spent waiting for outstanding GPU work before recycling a block (not allocator time)
double device_free_wait_seconds();

// Owning device allocation of doubles (or bytes via count*8).
struct DBuf {
  double* p = nullptr;
  size_t n = 0;
  DBuf() = default;
  explicit DBuf(size_t count) { alloc(count); }
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  DBuf(DBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DBuf& operator=(DBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DBuf() { release(); }
  void alloc(size_t count) {
    release();
    if (count == 0) return;
    p = static_cast<double*>(device_alloc(count * sizeof(double)));
    n = count;
  }
  void ensure(size_t count) { if (count > n) alloc(count); }
  void release() { if (p) device_free(p, n * sizeof(double)); p = nullptr; n = 0; }
  void zero(cudaStream_t s) { if (p) XTPB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(double), s)); }
};

inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

}  // namespace xtpb
