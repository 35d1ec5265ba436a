// NCCL plumbing for z-slab sharding.  NCCL is resolved at run time (dlopen of the libnccl.so.2 torch has
// already loaded), so libddpm3d.so has no link-time dependency on it and still loads on a CPU-only box.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace ddpm3d {

constexpr int SLAB_MAX_RANKS = 16;
constexpr int SLAB_GATHER_DOUBLES = 512;  // per rank and parity: B * 32 groups * 2 sums, B <= 8

struct SlabComm {
  void* comm = nullptr;  // ncclComm_t
  int rank = 0, world = 1;
  int z_begin = 0, z_total = 0;  // this rank's slab within the global volume (set per problem)
  bool enabled = false;          // set_slab(z_begin, z_total > 0) switches the sharded path on, set_slab(0, 0) off
  bool active() const { return comm != nullptr && world > 1 && enabled; }

  // ---- peer-mapped memory over NVLink (CUDA IPC), see comm.cu "peer path" -----------------------------------
  // Every rank exports a small mailbox (flags + GroupNorm statistics slots) to all ranks and its activation
  // workspace to its two z-neighbours.  GroupNorm / pack kernels then store their boundary planes straight into the
  // neighbour's halo planes and the statistics into every rank's mailbox; ordering is carried by sequence-numbered
  // flags, so a halo exchange costs two one-thread kernels instead of an NCCL send/recv group and the statistics
  // all-gather one small kernel instead of an ncclAllGather.
  bool p2p = false;                        // mailboxes mapped
  void* mailbox = nullptr;                 // mine (cudaMalloc): uint32 flags[256] | double gather[2][world][SLAB_GATHER_DOUBLES]
  void* peer_mailbox[SLAB_MAX_RANKS] = {}; // every rank's mailbox in this process' address space ([rank] = mine)
  char* peer_ws[2] = {nullptr, nullptr};   // workspace of rank - 1 / rank + 1
  unsigned char peer_ws_handle[2][64] = {};
  bool peer_ws_open[2] = {false, false};
  void* ipc_stage = nullptr;               // device staging for the handle all-gather
  // the exchange sequence numbers live in the mailbox (device side, incremented by the flag kernels themselves), so a
  // sharded step can be captured in a CUDA graph and replayed
  bool halo_p2p() const { return p2p && (rank == 0 || peer_ws[0]) && (rank + 1 == world || peer_ws[1]); }
};
// mailbox layout
constexpr size_t SLAB_FLAGS_BYTES = 4096;
constexpr int SLAB_F_READY_UP = 0, SLAB_F_READY_DOWN = 1, SLAB_F_CONSUMED_UP = 2, SLAB_F_CONSUMED_DOWN = 3, SLAB_F_STATS = 8;
constexpr int SLAB_F_HALO_SEQ = 64, SLAB_F_STATS_SEQ = 65;  // local counters (never written by peers)
inline size_t slab_mailbox_bytes() { return SLAB_FLAGS_BYTES + (size_t)2 * SLAB_MAX_RANKS * SLAB_GATHER_DOUBLES * sizeof(double); }

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


// ---- peer path ---------------------------------------------------------------------------------------------------
// collective (every rank of the communicator): allocate + exchange + map the mailboxes
int comm_peer_init(SlabComm* c);
// collective: make sure the z-neighbours' workspaces (`ws`, same size on every rank) are mapped; re-maps after a
// re-allocation anywhere (*changed is set: captured graphs hold the old addresses).  Synchronises the stream.
int comm_peer_sync_ws(SlabComm* c, void* ws, cudaStream_t s, bool* changed);
void comm_peer_destroy(SlabComm* c);
// halo handshake, one-thread kernels.  pre: start exchange seq = ++halo_seq; tell the neighbours that every halo up to
// seq - 1 has been consumed (all earlier kernels of this stream are done) and wait until they have consumed theirs --
// after it this rank may store into their halo planes.  post: tell the neighbours that the boundary planes of exchange
// seq are in their halo planes and wait for theirs.
int comm_halo_pre(const SlabComm& c, cudaStream_t s);
int comm_halo_post(const SlabComm& c, cudaStream_t s);
// statistics: seq = ++stats_seq; store `count` doubles into slot [seq & 1][rank] of every rank's mailbox and raise this
// rank's flag there.  The finalize kernel reads stats_seq, waits for all `world` flags and sums the slots.
int comm_stats_push(const SlabComm& c, const double* sums, int count, cudaStream_t s);
const double* comm_stats_slots(const SlabComm& c);        // [2][SLAB_MAX_RANKS][SLAB_GATHER_DOUBLES]
const uint32_t* comm_stats_flags(const SlabComm& c);      // [world]
const uint32_t* comm_stats_seq(const SlabComm& c);

}  // namespace ddpm3d
