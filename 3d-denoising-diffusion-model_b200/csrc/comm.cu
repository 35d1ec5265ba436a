#include "comm.h"

#include <dlfcn.h>
#include <stdint.h>
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
  comm_peer_destroy(c);
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


// =================================================================================================================
// peer path: CUDA IPC mappings + flag kernels
// =================================================================================================================
namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// a protocol bug (or a dead peer) must surface as a launch failure, never as a hung GPU: give up after ~4 s
__device__ __forceinline__ void spin_until(const uint32_t* flag, uint32_t seq, int what) {
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(flag) - seq) < 0) {
    if (clock64() - t0 > 8000000000LL) {
      printf("ddpm3d slab: peer flag %d timed out waiting for sequence %u (have %u)\n", what, seq, ld_acquire_sys(flag));
      __trap();
    }
  }
}

// flags: mine; up / down: the neighbours' flag arrays (NULL at the volume ends)
__global__ void halo_pre_kernel(uint32_t* flags, uint32_t* up, uint32_t* down) {
  if (threadIdx.x != 0) return;
  const uint32_t seq = flags[SLAB_F_HALO_SEQ] + 1;
  flags[SLAB_F_HALO_SEQ] = seq;
  __threadfence_system();
  if (up) st_release_sys(up + SLAB_F_CONSUMED_DOWN, seq - 1);   // I am its lower neighbour
  if (down) st_release_sys(down + SLAB_F_CONSUMED_UP, seq - 1);
  if (up) spin_until(flags + SLAB_F_CONSUMED_UP, seq - 1, SLAB_F_CONSUMED_UP);
  if (down) spin_until(flags + SLAB_F_CONSUMED_DOWN, seq - 1, SLAB_F_CONSUMED_DOWN);
}
__global__ void halo_post_kernel(uint32_t* flags, uint32_t* up, uint32_t* down) {
  if (threadIdx.x != 0) return;
  const uint32_t seq = flags[SLAB_F_HALO_SEQ];
  __threadfence_system();  // (the producing kernel has completed: its peer stores are performed; this orders the flag after them)
  if (up) st_release_sys(up + SLAB_F_READY_DOWN, seq);
  if (down) st_release_sys(down + SLAB_F_READY_UP, seq);
  if (up) spin_until(flags + SLAB_F_READY_UP, seq, SLAB_F_READY_UP);
  if (down) spin_until(flags + SLAB_F_READY_DOWN, seq, SLAB_F_READY_DOWN);
}

struct PushArgs {
  void* box[SLAB_MAX_RANKS];
  int world, rank, count;
};
// one block: copy the sums into every rank's slot, fence, raise this rank's flag everywhere
__global__ void __launch_bounds__(256) stats_push_kernel(const double* __restrict__ sums, PushArgs a) {
  uint32_t* mine = reinterpret_cast<uint32_t*>(a.box[a.rank]);
  const uint32_t seq = mine[SLAB_F_STATS_SEQ] + 1;
  for (int r = 0; r < a.world; ++r) {
    double* dst = reinterpret_cast<double*>((char*)a.box[r] + SLAB_FLAGS_BYTES) +
                  ((size_t)(seq & 1u) * SLAB_MAX_RANKS + a.rank) * SLAB_GATHER_DOUBLES;
    for (int i = threadIdx.x; i < a.count; i += blockDim.x) dst[i] = sums[i];
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < a.world) st_release_sys(reinterpret_cast<uint32_t*>(a.box[threadIdx.x]) + SLAB_F_STATS + a.rank, seq);
  if (threadIdx.x == 0) mine[SLAB_F_STATS_SEQ] = seq;
}

int allgather_bytes(SlabComm* c, const void* host_in, void* host_out, size_t bytes, cudaStream_t s) {
  char* stage = (char*)c->ipc_stage;  // [bytes | world * bytes]
  DD_CUDA(cudaMemcpyAsync(stage, host_in, bytes, cudaMemcpyHostToDevice, s));
  DD_NCCL(api().AllGather(stage, stage + 256, bytes, NCCL_INT8, (NcclComm)c->comm, s));
  DD_CUDA(cudaMemcpyAsync(host_out, stage + 256, bytes * c->world, cudaMemcpyDeviceToHost, s));
  DD_CUDA(cudaStreamSynchronize(s));
  return DDPM3D_OK;
}

}  // namespace

int comm_peer_init(SlabComm* c) {
  DD_CHECK(c->comm && c->world <= SLAB_MAX_RANKS, DDPM3D_ERR_STATE, "peer path: communicator missing or too many ranks");
  comm_peer_destroy(c);
  DD_CUDA(cudaMalloc(&c->ipc_stage, 256 + 64 * SLAB_MAX_RANKS));
  DD_CUDA(cudaMalloc(&c->mailbox, slab_mailbox_bytes()));
  DD_CUDA(cudaMemset(c->mailbox, 0, slab_mailbox_bytes()));
  DD_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t mine;
  DD_CUDA(cudaIpcGetMemHandle(&mine, c->mailbox));
  unsigned char all[64 * SLAB_MAX_RANKS];
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  DD_TRY(allgather_bytes(c, &mine, all, 64, nullptr));
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) { c->peer_mailbox[r] = c->mailbox; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, all + 64 * r, 64);
    const cudaError_t e = cudaIpcOpenMemHandle(&c->peer_mailbox[r], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {  // no peer access between these devices: stay on the NCCL path
      cudaGetLastError();
      for (int q = 0; q < r; ++q)
        if (q != c->rank && c->peer_mailbox[q]) cudaIpcCloseMemHandle(c->peer_mailbox[q]);
      memset(c->peer_mailbox, 0, sizeof(c->peer_mailbox));
      c->p2p = false;
      return DDPM3D_OK;
    }
  }
  c->p2p = true;
  return DDPM3D_OK;
}

int comm_peer_sync_ws(SlabComm* c, void* ws, cudaStream_t s, bool* changed) {
  *changed = false;
  if (!c->p2p) return DDPM3D_OK;
  cudaIpcMemHandle_t mine;
  DD_CUDA(cudaIpcGetMemHandle(&mine, ws));
  unsigned char all[64 * SLAB_MAX_RANKS];
  DD_TRY(allgather_bytes(c, &mine, all, 64, s));
  const int nb[2] = {c->rank - 1, c->rank + 1};
  for (int d = 0; d < 2; ++d) {
    if (nb[d] < 0 || nb[d] >= c->world) continue;
    const unsigned char* h = all + 64 * nb[d];
    if (c->peer_ws_open[d] && memcmp(h, c->peer_ws_handle[d], 64) == 0) continue;
    *changed = true;
    if (c->peer_ws_open[d]) {
      cudaIpcCloseMemHandle(c->peer_ws[d]);
      c->peer_ws_open[d] = false;
      c->peer_ws[d] = nullptr;
    }
    cudaIpcMemHandle_t hh;
    memcpy(&hh, h, 64);
    void* p = nullptr;
    DD_CUDA(cudaIpcOpenMemHandle(&p, hh, cudaIpcMemLazyEnablePeerAccess));
    c->peer_ws[d] = (char*)p;
    c->peer_ws_open[d] = true;
    memcpy(c->peer_ws_handle[d], h, 64);
  }
  return DDPM3D_OK;
}

void comm_peer_destroy(SlabComm* c) {
  for (int d = 0; d < 2; ++d) {
    if (c->peer_ws_open[d]) cudaIpcCloseMemHandle(c->peer_ws[d]);
    c->peer_ws_open[d] = false;
    c->peer_ws[d] = nullptr;
  }
  for (int r = 0; r < SLAB_MAX_RANKS; ++r) {
    if (c->peer_mailbox[r] && c->peer_mailbox[r] != c->mailbox) cudaIpcCloseMemHandle(c->peer_mailbox[r]);
    c->peer_mailbox[r] = nullptr;
  }
  if (c->mailbox) cudaFree(c->mailbox);
  if (c->ipc_stage) cudaFree(c->ipc_stage);
  c->mailbox = c->ipc_stage = nullptr;
  c->p2p = false;
}

int comm_halo_pre(const SlabComm& c, cudaStream_t s) {
  uint32_t* up = c.rank > 0 ? (uint32_t*)c.peer_mailbox[c.rank - 1] : nullptr;
  uint32_t* down = c.rank + 1 < c.world ? (uint32_t*)c.peer_mailbox[c.rank + 1] : nullptr;
  halo_pre_kernel<<<1, 32, 0, s>>>((uint32_t*)c.mailbox, up, down);
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

int comm_halo_post(const SlabComm& c, cudaStream_t s) {
  uint32_t* up = c.rank > 0 ? (uint32_t*)c.peer_mailbox[c.rank - 1] : nullptr;
  uint32_t* down = c.rank + 1 < c.world ? (uint32_t*)c.peer_mailbox[c.rank + 1] : nullptr;
  halo_post_kernel<<<1, 32, 0, s>>>((uint32_t*)c.mailbox, up, down);
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

int comm_stats_push(const SlabComm& c, const double* sums, int count, cudaStream_t s) {
  DD_CHECK(count <= SLAB_GATHER_DOUBLES, DDPM3D_ERR_ARG, "peer path: statistics record too large (batch > 8)");
  PushArgs a{};
  for (int r = 0; r < c.world; ++r) a.box[r] = c.peer_mailbox[r];
  a.world = c.world; a.rank = c.rank; a.count = count;
  stats_push_kernel<<<1, 256, 0, s>>>(sums, a);
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

const double* comm_stats_slots(const SlabComm& c) { return reinterpret_cast<const double*>((const char*)c.mailbox + SLAB_FLAGS_BYTES); }
const uint32_t* comm_stats_flags(const SlabComm& c) { return reinterpret_cast<const uint32_t*>(c.mailbox) + SLAB_F_STATS; }
const uint32_t* comm_stats_seq(const SlabComm& c) { return reinterpret_cast<const uint32_t*>(c.mailbox) + SLAB_F_STATS_SEQ; }

}  // namespace ddpm3d
