// Context, error reporting, dense solver wrappers (cuSOLVER eigh / inverse; these are the
// "serial-ish library calls" of SURVEY.md section 7 hard part 6 and are timed apart from the contractions).
#include "internal.h"

namespace xtpb {

namespace {
thread_local std::string g_last_error;
}
void set_last_error(const std::string& msg) { g_last_error = msg; }
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
