// Probe for the Sigma_c grid-scan kernels: how fast can B200 (sm_100a) produce FP64
// reciprocals?  Measures (a) MUFU.RCP64H alone (rcp.approx.ftz.f64, ILP independent chains per thread), (b) the full
// rcp_fast sequence of csrc/kernels.cu (seed + 3 DFMA), (c) the damped-kernel polynomial (14 FP64 operations, no SFU),
// (d) the latency of a dependent DFMA chain.  One JSON object on stdout.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_rcp64 tools/probe_rcp64.cu && tools/probe_rcp64
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s @%d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ double rcp_seed(double x) {
  double r;
  asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
}
__device__ __forceinline__ double rcp_fast(double x) {
  const double r = rcp_seed(x);
  const double e = fma(-x, r, 1.0);
  const double t = fma(e, e, e);
  return fma(r, t, r);
}
// alternative seed: FP32 MUFU.RCP on the rounded argument, two Newton steps in FP64 (4 DFMA + 2 conversions)
__device__ __forceinline__ double rcp_f32seed(double x) {
  float rf;
  asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"((float)x));
  double r = (double)rf;
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}
__device__ __forceinline__ double damped_poly(double x) {
  const double u = x * x;
  double h = 4.7076855031461505630e+01;
  h = fma(h, u, -1.9377648362907347975e+02); h = fma(h, u, 6.7736136257549954211e+02);
  h = fma(h, u, -1.9817217134061364861e+03); h = fma(h, u, 4.7687717542357461594e+03);
  h = fma(h, u, -9.2407715743593665354e+03); h = fma(h, u, 1.4044288705238401560e+04);
  h = fma(h, u, -1.6186442488460224111e+04); h = fma(h, u, 1.3530243473087691496e+04);
  h = fma(h, u, -7.7113140956002125264e+03); h = fma(h, u, 2.7346181506141992876e+03);
  h = fma(h, u, -5.1951515218134633193e+02); h = fma(h, u, 3.9478417604357434475e+01);
  return x * h;
}

// MODE 0: seed only, 1: rcp_fast, 2: damped polynomial.  Each thread keeps ILP independent values x_i <- f(x_i) + c.
template <int MODE, int ILP>
__global__ void __launch_bounds__(1024) unary_loop(double* out, int iters, double c) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = 1.5 + 1e-3 * threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      const double v = MODE == 0 ? rcp_seed(x[i])
                     : MODE == 1 ? rcp_fast(x[i])
                     : MODE == 2 ? damped_poly(x[i] * 0.01)
                     : MODE == 3 ? rcp_f32seed(x[i])
                                 : fma(fma(fma(fma(x[i], 0.99, 0.1), 0.98, 0.2), 0.97, 0.3), 0.96, 0.4);   // 4 DFMA
      x[i] = v + c;              // one DADD per evaluation, like the subtraction w - z of the kernel
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 123.456) out[0] = s;
}

// dependent DFMA chain: one warp per SM, clock64 around `iters` dependent operations
__global__ void dfma_latency(double* out, long long* cycles, int iters, double a, double b) {
  double x = 1.0 + threadIdx.x * 1e-9;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) x = fma(x, a, b);
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  if (x == 123.456) out[0] = x;
}

template <typename F> float time_ms(F f, int rep = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < rep; ++r) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
  }
  return best;
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  double* out; long long* cyc; CK(cudaMalloc(&out, 8)); CK(cudaMalloc(&cyc, 8));
  printf("{\"gpu\":\"%s\",\"sms\":%d,\"clock_khz\":%d,\n\"rates\":[", p.name, sms, khz);
  const int iters = 4000;
  bool first = true;
  auto report = [&](const char* what, int warps, int ilp, float ms) {
    const double evals = (double)iters * ilp * warps * 32.0 * sms;
    const double per_clk_sm = evals / (ms * 1e-3) / ((double)khz * 1e3) / sms;
    printf("%s{\"op\":\"%s\",\"warps_per_sm\":%d,\"ilp\":%d,\"ms\":%.4f,\"gevals_per_s\":%.1f,\"per_clk_per_sm\":%.2f}",
           first ? "" : ",\n", what, warps, ilp, ms, evals / ms * 1e-6, per_clk_sm);
    first = false;
  };
  for (int warps : {8, 16, 32}) {
    report("mufu_rcp64h+dadd", warps, 8, time_ms([&] { unary_loop<0, 8><<<sms, warps * 32>>>(out, iters, 0.25); }));
    report("rcp_fast+dadd", warps, 8, time_ms([&] { unary_loop<1, 8><<<sms, warps * 32>>>(out, iters, 0.25); }));
    report("damped_poly+dadd", warps, 8, time_ms([&] { unary_loop<2, 8><<<sms, warps * 32>>>(out, iters, 0.25); }));
    report("rcp_f32seed+dadd", warps, 8, time_ms([&] { unary_loop<3, 8><<<sms, warps * 32>>>(out, iters, 0.25); }));
    report("4dfma+dadd", warps, 8, time_ms([&] { unary_loop<4, 8><<<sms, warps * 32>>>(out, iters, 0.25); }));
  }
  report("rcp_fast+dadd", 16, 2, time_ms([&] { unary_loop<1, 2><<<sms, 16 * 32>>>(out, iters, 0.25); }));
  report("rcp_fast+dadd", 16, 4, time_ms([&] { unary_loop<1, 4><<<sms, 16 * 32>>>(out, iters, 0.25); }));
  const int lat_iters = 100000;
  dfma_latency<<<1, 32>>>(out, cyc, lat_iters, 1.0000001, 1e-9);
  CK(cudaDeviceSynchronize());
  long long c = 0; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
  printf("],\n\"dfma_dependent_latency_cycles\":%.2f}\n", (double)c / lat_iters);
  return 0;
}
