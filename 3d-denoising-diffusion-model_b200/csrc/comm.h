// NCCL plumbing for z-slab sharding.  NCCL is resolved at run time (dlopen of the libnccl.so.2 torch has
// already loaded), so libddpm3d.so has no link-time dependency on it and still loads on a CPU-only box.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>

namespace ddpm3d {

struct SlabComm {
  void* comm = nullptr;  // ncclComm_t
  int rank = 0, world = 1;
  int z_begin = 0, z_total = 0;  // this rank's slab within the global volume (set per problem)
  bool enabled = false;          // set_slab(z_begin, z_total > 0) switches the sharded path on, set_slab(0, 0) off
  bool active() const { return comm != nullptr && world > 1 && enabled; }
};

int comm_unique_id(void* out128);
int comm_init(SlabComm* c, const void* id128, int rank, int world);
void comm_destroy(SlabComm* c);
// exchange the boundary planes of a [B][Zl+2][plane_bytes] tensor with the z-neighbours; volume ends are zeroed
int comm_halo_exchange(const SlabComm& c, void* base, int B, int Zl, size_t plane_bytes, cudaStream_t s);
// all-gather `count` doubles per rank: recv[world][count]
int comm_allgather_f64(const SlabComm& c, const double* send, double* recv, size_t count, cudaStream_t s);
// all-gather `bytes` bytes per rank and per batch element: recv + b * recv_bstride = [world][bytes] (rank order = z order)
int comm_allgather_slabs(const SlabComm& c, const void* send, void* recv, int B, size_t bytes, size_t send_bstride,
                         size_t recv_bstride, cudaStream_t s);

}  // namespace ddpm3d
