// Multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.  The reference has no distributed layer on
// this path (one process, an OpenMP thread per GPU, host-side reductions -- upstream xtp/src/libxtp/openmp_cuda.cc);
// here every rank owns a cyclic share of the second tensor index and the partial results of each stage are
// all-reduced on the device.  The unique id is created by rank 0 (xtpb_comm_unique_id) and
// distributed by the host program (torch.distributed / MPI / a file) before xtpb_ctx_comm_init.
//
// NCCL is bound at run time (dlopen), not at link time: a process that has already loaded a libnccl.so.2 (PyTorch
// ships its own) must keep using that copy, and a single-GPU process must not need NCCL at all.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "internal.h"

namespace xtpb {

namespace {
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi& nccl_api() {
  static NcclApi api;
  static bool loaded = false;
  if (loaded) return api;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);     // the copy already in the process
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) throw Error(std::string("xtpb: cannot load libnccl.so.2 (needed for world > 1): ") + dlerror());
  auto sym = [&](const char* name) {
    void* p = dlsym(h, name);
    if (!p) throw Error(std::string("xtpb: libnccl.so.2 lacks ") + name);
    return p;
  };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.Reduce = reinterpret_cast<decltype(api.Reduce)>(sym("ncclReduce"));
  api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  loaded = true;
  return api;
}
}  // namespace

#define XTPB_NCCL(expr)                                                                               \
  do {                                                                                                \
    ncclResult_t r__ = (expr);                                                                        \
    if (r__ != ncclSuccess)                                                                           \
      throw ::xtpb::Error(std::string("NCCL error ") + nccl_api().GetErrorString(r__) + " in " #expr); \
  } while (0)

static_assert(sizeof(ncclUniqueId) == 128, "xtpb_comm_unique_id hands out 128 bytes");

void comm_unique_id(char* out_128) {
  ncclUniqueId id;
  XTPB_NCCL(nccl_api().GetUniqueId(&id));
  std::memcpy(out_128, &id, sizeof(id));
}

void Context::comm_init(const char* unique_id_128, int rank_, int world_) {
  XTPB_REQUIRE(world_ >= 1 && rank_ >= 0 && rank_ < world_, "bad rank / world size");
  XTPB_REQUIRE(nccl == nullptr, "communicator already initialised");
  if (world_ == 1) return;
  ncclUniqueId id;
  std::memcpy(&id, unique_id_128, sizeof(id));
  ncclComm_t c = nullptr;
  XTPB_CUDA(cudaSetDevice(device));
  XTPB_NCCL(nccl_api().CommInitRank(&c, world_, id, rank_));
  nccl = c;
  rank = rank_;
  world = world_;
  XTPB_CUDA(cudaStreamCreateWithFlags(&comm_stream, cudaStreamNonBlocking));
}

void Context::comm_destroy() {
  if (nccl) {
    cudaDeviceSynchronize();
    nccl_api().CommDestroy(static_cast<ncclComm_t>(nccl));
    nccl = nullptr;
  }
  if (comm_stream) {
    cudaStreamDestroy(comm_stream);
    comm_stream = nullptr;
  }
  rank = 0;
  world = 1;
}

void Context::allreduce_sum(double* buf, size_t count, cudaStream_t st) {
  if (world == 1 || count == 0) return;
  cudaStream_t q = st ? st : stream;
  const int slot = prof_begin(PROF_COMM, 8.0 * (double)count, q);     // work = payload bytes
  XTPB_NCCL(nccl_api().AllReduce(buf, buf, count, ncclDouble, ncclSum, static_cast<ncclComm_t>(nccl), q));
  prof_end(slot, q);
}

// sum over ranks delivered to `root` only (in place there; the other ranks' buffers are left as they were)
void Context::reduce_sum(double* buf, size_t count, int root, cudaStream_t st) {
  if (world == 1 || count == 0) return;
  cudaStream_t q = st ? st : stream;
  const int slot = prof_begin(PROF_COMM, 8.0 * (double)count, q);
  XTPB_NCCL(nccl_api().Reduce(buf, buf, count, ncclDouble, ncclSum, root, static_cast<ncclComm_t>(nccl), q));
  prof_end(slot, q);
}
void Context::bcast(double* buf, size_t count, int root, cudaStream_t st) {
  if (world == 1 || count == 0) return;
  cudaStream_t q = st ? st : stream;
  const int slot = prof_begin(PROF_COMM, 8.0 * (double)count, q);
  XTPB_NCCL(nccl_api().Broadcast(buf, buf, count, ncclDouble, root, static_cast<ncclComm_t>(nccl), q));
  prof_end(slot, q);
}
void Context::group_start() { if (world > 1) XTPB_NCCL(nccl_api().GroupStart()); }
void Context::group_end() { if (world > 1) XTPB_NCCL(nccl_api().GroupEnd()); }

void Context::allgather(const double* send, double* recv, size_t count_per_rank, cudaStream_t st) {
  if (world == 1) {
    if (send != recv)
      XTPB_CUDA(cudaMemcpyAsync(recv, send, count_per_rank * 8, cudaMemcpyDeviceToDevice, st ? st : stream));
    return;
  }
  cudaStream_t q = st ? st : stream;
  const int slot = prof_begin(PROF_COMM, 8.0 * (double)count_per_rank * world, q);
  XTPB_NCCL(nccl_api().AllGather(send, recv, count_per_rank, ncclDouble, static_cast<ncclComm_t>(nccl), q));
  prof_end(slot, q);
}

}  // namespace xtpb
