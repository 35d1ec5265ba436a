#include "comm.h"

#include <dlfcn.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace ddpm3d {

namespace {

// the slice of nccl.h this file needs (ABI-stable since NCCL 2.x)
typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
enum { NCCL_INT8 = 0, NCCL_FLOAT64 = 8 };

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.lib) break;
    }
    if (!a.lib) return;
#define LOAD(field, sym) *(void**)(&a.field) = dlsym(a.lib, sym)
    LOAD(GetUniqueId, "ncclGetUniqueId");
    LOAD(CommInitRank, "ncclCommInitRank");
    LOAD(CommDestroy, "ncclCommDestroy");
    LOAD(Send, "ncclSend");
    LOAD(Recv, "ncclRecv");
    LOAD(AllGather, "ncclAllGather");
    LOAD(GroupStart, "ncclGroupStart");
    LOAD(GroupEnd, "ncclGroupEnd");
    LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.Send && a.Recv && a.AllGather && a.GroupStart && a.GroupEnd;
  });
  return a;
}

#define DD_NCCL(expr)                                                                                          \
  do {                                                                                                         \
    const int r__ = (expr);                                                                                    \
    if (r__ != 0) {                                                                                            \
      set_error(std::string(#expr) + ": " + (api().GetErrorString ? api().GetErrorString(r__) : "nccl error")); \
      return DDPM3D_ERR_CUDA;                                                                                  \
    }                                                                                                          \
  } while (0)

}  // namespace

int comm_unique_id(void* out128) {
  DD_CHECK(api().ok, DDPM3D_ERR_CUDA, "NCCL (libnccl.so.2) could not be loaded");
  NcclUniqueId id;
  DD_NCCL(api().GetUniqueId(&id));
  memcpy(out128, &id, 128);
  return DDPM3D_OK;
}

int comm_init(SlabComm* c, const void* id128, int rank, int world) {
  DD_CHECK(api().ok, DDPM3D_ERR_CUDA, "NCCL (libnccl.so.2) could not be loaded");
  DD_CHECK(world >= 1 && rank >= 0 && rank < world, DDPM3D_ERR_ARG, "set_comm: bad rank / world");
  comm_destroy(c);
  NcclUniqueId id;
  memcpy(&id, id128, 128);
  NcclComm comm = nullptr;
  DD_NCCL(api().CommInitRank(&comm, world, id, rank));
  c->comm = comm;
  c->rank = rank;
  c->world = world;
  return DDPM3D_OK;
}

void comm_destroy(SlabComm* c) {
  if (c->comm && api().ok) api().CommDestroy((NcclComm)c->comm);
  c->comm = nullptr;
  c->world = 1;
  c->rank = 0;
  c->enabled = false;
}

int comm_halo_exchange(const SlabComm& c, void* base, int B, int Zl, size_t plane_bytes, cudaStream_t s) {
  char* p = (char*)base;
  const size_t bstride = (size_t)(Zl + 2) * plane_bytes;
  const bool has_up = c.rank > 0, has_down = c.rank + 1 < c.world;
  for (int b = 0; b < B; ++b) {
    if (!has_up) DD_CUDA(cudaMemsetAsync(p + b * bstride, 0, plane_bytes, s));
    if (!has_down) DD_CUDA(cudaMemsetAsync(p + b * bstride + (size_t)(Zl + 1) * plane_bytes, 0, plane_bytes, s));
  }
  if (!has_up && !has_down) return DDPM3D_OK;
  NcclComm comm = (NcclComm)c.comm;
  DD_NCCL(api().GroupStart());
  for (int b = 0; b < B; ++b) {
    char* t = p + b * bstride;
    if (has_up) {  // my first real plane -> upper neighbour's trailing halo; its last real plane -> my leading halo
      DD_NCCL(api().Send(t + plane_bytes, plane_bytes, NCCL_INT8, c.rank - 1, comm, s));
      DD_NCCL(api().Recv(t, plane_bytes, NCCL_INT8, c.rank - 1, comm, s));
    }
    if (has_down) {
      DD_NCCL(api().Send(t + (size_t)Zl * plane_bytes, plane_bytes, NCCL_INT8, c.rank + 1, comm, s));
      DD_NCCL(api().Recv(t + (size_t)(Zl + 1) * plane_bytes, plane_bytes, NCCL_INT8, c.rank + 1, comm, s));
    }
  }
  DD_NCCL(api().GroupEnd());
  return DDPM3D_OK;
}

int comm_allgather_f64(const SlabComm& c, const double* send, double* recv, size_t count, cudaStream_t s) {
  DD_NCCL(api().AllGather(send, recv, count, NCCL_FLOAT64, (NcclComm)c.comm, s));
  return DDPM3D_OK;
}

int comm_allgather_slabs(const SlabComm& c, const void* send, void* recv, int B, size_t bytes, size_t send_bstride,
                         size_t recv_bstride, cudaStream_t s) {
  DD_NCCL(api().GroupStart());
  for (int b = 0; b < B; ++b)
    DD_NCCL(api().AllGather((const char*)send + b * send_bstride, (char*)recv + b * recv_bstride, bytes, NCCL_INT8, (NcclComm)c.comm, s));
  DD_NCCL(api().GroupEnd());
  return DDPM3D_OK;
}

}  // namespace ddpm3d
