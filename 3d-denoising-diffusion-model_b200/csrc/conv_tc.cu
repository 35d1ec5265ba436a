// 3x3x3 / 1x1x1 convolution as an implicit GEMM on the 5th-generation tensor cores (sm_100a).
//
//   D[m, n] = sum_k A[m, k] * W[n, k]      m = output voxel, n = output channel,
//                                          k = (tap, input channel) [+ channels of 1x1x1 sources]
//
// * A is never materialised: every k-step (one tap, 64 channels) is ONE 5-D TMA box
//   {64 ch, bw, bh, bz, 1} of the channels-last activation tensor, shifted by the tap offset.
//   Out-of-bounds coordinates (the conv's zero padding, and bricks overhanging the volume) are
//   zero-filled by the TMA unit, so there is no halo handling in the kernel.
// * The box lands in shared memory as 128 rows x 128 bytes with the 128-byte swizzle, which is
//   exactly the canonical K-major SWIZZLE_128B operand layout of tcgen05.mma.
// * W ([Cout][Ktot] bf16, K contiguous) comes from a 2-D TMA box {64, BN}.
// * Accumulators live in TMEM (2 x BN fp32 columns, double buffered): the epilogue of tile i
//   (tcgen05.ld -> +bias -> +residual (plain / 2x2-average / nearest-upsampled) -> bf16 -> global)
//   overlaps the MMAs of tile i+1.
// * Persistent CTAs (one per SM), warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer
//   (+ TMEM allocation), warps 2-5 = epilogue.  smem ring of NSTAGE {A,B} slots with full/empty
//   mbarriers; tcgen05.commit releases slots and publishes finished accumulators.
//
// The ResBlock skip_connection (1x1x1 conv over the un-normalised block input, which for decoder
// blocks is a channel concat of two tensors) is folded in as extra k-steps reading those tensors
// in place (unet.py:222,254-256 and :1040-1042).
#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <type_traits>

#include "kernels.h"
#include "tc_ptx.cuh"

namespace ddpm3d {

namespace {

constexpr int BM = 128;      // UMMA M (rows of the brick, <= 128 valid)
constexpr int NTHREADS = 192;
constexpr int A_BYTES = BM * BK * 2;  // 16 KB
constexpr int CS_TR_BYTES = 4 * 32 * 33 * 4;  // transpose scratch of the four epilogue warps
constexpr int CS_SMEM_MAX = CS_TR_BYTES + 16 * 1024;  // + channel-sum accumulators: 4 warps x B x Cout x 2 floats

struct TcParams {
  int nsrc;
  int chunks[3];     // C/64 of each source
  int n_main_steps;  // taps * chunks[0]
  int nk;            // total k-steps
  int taps;          // 27 or 1
  int zoff;          // halo planes in front of the main source (z-slab sharding)
  int stride;        // (1, s, s) stride of the main source (Downsample conv, unet.py:129-133); extra sources have none
  int bw, bh, bz;    // brick (one 128-row MMA tile)
  int pw, ph, pz;    // offset of the second brick of a CTA tile (MT == 2): exactly one is non-zero
  int nWt, nHt, nZt, nNt;  // CTA tiles per dimension (a CTA tile = MT bricks)
  int num_tiles;
  int B, Z, Ho, Wo, Cout;
  uint32_t a_tx_bytes;  // bytes one A box delivers
  const float* bias;
  const void* res;   // T
  int res_mode;
  void* out;         // T
  // stream-K (layers with too few tiles to fill the GPU): the num_tiles * nk k-steps are dealt evenly to the CTAs;
  // a CTA that owns only part of a tile's K range writes its fp32 accumulators to
  // partial[tile][slot][MT*128][BN] (slot = position among the tile's contributors) and a fix-up kernel adds
  // the slots in order and applies bias / residual.  Tiles owned by one CTA take the normal epilogue.
  int sk;
  int max_slots;
  float* partial;
  float* chsum;      // [B][cs_slots][Cout][2] per-CTA channel sums of the output, or NULL
  int cs_slots;      // = chsum_slots() (one slot per SM)
  uint32_t cs_off;   // byte offset of the channel-sum accumulators in dynamic smem
};

template <typename T>
__device__ __forceinline__ void add8r(float* v, const uint4& raw) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float a, b;
    unpack2<T>(w[i], a, b);
    v[2 * i] += a;
    v[2 * i + 1] += b;
  }
}
template <typename T>
__device__ __forceinline__ void add8(float* v, const T* p, float scale) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float a, b;
    unpack2<T>(w[i], a, b);
    v[2 * i] += scale * a;
    v[2 * i + 1] += scale * b;
  }
}

// the sequence of (tile, k-range) segments of one CTA; identical in the three warp roles
// With CL == 2 the two CTAs of a cluster walk the same sequence of super-tiles in lock step: both take the same
// output-channel tile (so the weight tile is fetched once and multicast) and adjacent M tiles; `tile` may be a
// dummy (>= num_tiles) for the odd one out, which loads zeros and stores nothing.
struct WorkIter {
  int sk, nk, num_tiles, nNt, cur, step, cl, crank;
  int64_t kpos, kend;
  __device__ WorkIter(const TcParams& p, int CL, int rank)
      : sk(p.sk), nk(p.nk), num_tiles(p.num_tiles), nNt(p.nNt), cur(blockIdx.x / CL), step(gridDim.x / CL), cl(CL), crank(rank) {
    const int64_t T = (int64_t)p.num_tiles * p.nk;
    kpos = (int64_t)blockIdx.x * T / gridDim.x;
    kend = (int64_t)(blockIdx.x + 1) * T / gridDim.x;
  }
  __device__ bool next(int& tile, int& k0, int& k1) {
    if (!sk) {
      const int numM = num_tiles / nNt;
      const int n_super = ((numM + cl - 1) / cl) * nNt;
      if (cur >= n_super) return false;
      const int mp = cur / nNt, nt = cur - mp * nNt;
      const int m = mp * cl + crank;
      tile = m < numM ? m * nNt + nt : num_tiles + nt;  // dummy keeps the right n-tile for the weight multicast
      k0 = 0; k1 = nk;
      cur += step;
      return true;
    }
    if (kpos >= kend) return false;
    tile = (int)(kpos / nk);
    k0 = (int)(kpos - (int64_t)tile * nk);
    k1 = (int)min((int64_t)nk, k0 + (kend - kpos));
    kpos += k1 - k0;
    return true;
  }
};
// first CTA whose k-range touches k-step x (ranges are [c*T/G, (c+1)*T/G))
__host__ __device__ inline int sk_owner(int64_t x, int64_t T, int G) { return (int)(((x + 1) * G - 1) / T); }

// T: format of the main source and its weight columns; TS: format of the extra (1x1x1) sources and their weight
// columns, of the residual and of the output (both 16 bit; the MMA instruction descriptor is chosen per k-step)
template <typename T, typename TS, int MT, int BN, int NSTAGE, int CL>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapW, const TcParams p) {
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = MT * A_BYTES + B_BYTES;
  constexpr int ACC_COLS = MT * BN;       // fp32 accumulator columns of one CTA tile
  constexpr int TMEM_COLS = 2 * ACC_COLS;  // double buffered; 128, 256 or 512: a power of two >= 32
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM allocation must be a power of two <= 512");
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * NSTAGE + 4];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms are 1024-byte aligned
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * NSTAGE + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * NSTAGE + 2 + a); };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA0);
    prefetch_tmap(&mapW);
    if (p.nsrc > 1) prefetch_tmap(&mapA1);
    if (p.nsrc > 2) prefetch_tmap(&mapA2);
    // a stage is refilled by this CTA's producer AND (weight half) by the peer's: it is free only when the MMA
    // warps of all CL CTAs have consumed it
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), CL); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    fence_barrier_init();
  }
  if (warp == 1) {  // TMEM allocation is warp-wide
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
  if (CL > 1) cluster_sync_all();  // the peer's barriers are initialised before any multicast can signal them
  // programmatic dependent launch: everything above overlapped the tail of the kernel before this one (the GroupNorm
  // apply pass).  The kernel after this one (the GroupNorm finalize) is released by the MMA warp once it has issued its
  // last tile -- released at the start, it and the apply grid behind it would sit in griddepcontrol.wait next to the
  // running convolution for its whole duration (measured: +3 % on the network step)
  pdl_wait();

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      WorkIter it(p, CL, crank);
      int tile, kbeg, kend;
      while (it.next(tile, kbeg, kend)) {
        const int nt = tile % p.nNt;
        int m = tile / p.nNt;
        const int wt = m % p.nWt; m /= p.nWt;
        const int ht = m % p.nHt; m /= p.nHt;
        const int zt = m % p.nZt;
        const int b = m / p.nZt;  // == p.B for a dummy tile: every box is out of bounds and arrives as zeros
        const int w0 = wt * p.bw * (p.pw ? MT : 1), h0 = ht * p.bh * (p.ph ? MT : 1), z0 = zt * p.bz * (p.pz ? MT : 1);
        const int n0 = nt * BN;
        for (int kk = kbeg; kk < kend; ++kk) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_expect_tx(full_bar(stage), MT * p.a_tx_bytes + B_BYTES);
          const uint32_t a_dst = smem_base + stage * STAGE_BYTES;
          const uint32_t b_dst = a_dst + MT * A_BYTES;
          const CUtensorMap* map;
          int c0, dz = 0, dh = 0, dw = 0;
          if (kk < p.n_main_steps) {
            int tap = p.taps > 1 ? kk / p.chunks[0] : 13;
            c0 = (kk - (p.taps > 1 ? tap * p.chunks[0] : 0)) * BK;
            if (p.taps == 9) tap += 9;  // a 3x3 kernel of a dims = 2 network: the centre z-plane of the 27-tap numbering
            dz = tap / 9 - 1 + p.zoff; dh = (tap / 3) % 3 - 1; dw = tap % 3 - 1;
            map = &mapA0;
          } else {
            const int e = kk - p.n_main_steps;
            if (e < p.chunks[1]) { map = &mapA1; c0 = e * BK; }
            else { map = &mapA2; c0 = (e - p.chunks[1]) * BK; }
          }
          const int st = kk < p.n_main_steps ? p.stride : 1;  // input voxel = stride * output voxel + tap offset
#pragma unroll
          for (int j = 0; j < MT; ++j)
            tma_load_5d(a_dst + j * A_BYTES, map, full_bar(stage), c0, (w0 + j * p.pw) * st + dw, (h0 + j * p.ph) * st + dh,
                        z0 + j * p.pz + dz, b);
          if (CL == 1) {
            tma_load_2d(b_dst, &mapW, full_bar(stage), kk * BK, n0);
          } else {  // this CTA fetches its half of the weight tile for the whole cluster
            constexpr int HALF = BN / CL;
            tma_load_2d_mc(b_dst + crank * HALF * BK * 2, &mapW, full_bar(stage), kk * BK, n0 + crank * HALF,
                           (uint16_t)((1u << CL) - 1));
          }
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      constexpr uint32_t idesc_main = make_idesc(BM, BN, !std::is_same<T, f16>::value);
      constexpr uint32_t idesc_extra = make_idesc(BM, BN, !std::is_same<TS, f16>::value);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      WorkIter it(p, CL, crank);
      int tile, kbeg, kend;
      while (it.next(tile, kbeg, kend)) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
        for (int kk = kbeg; kk < kend; ++kk) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * STAGE_BYTES;
          const uint64_t bdesc = make_sw128_desc(a_addr + MT * A_BYTES);
          const uint32_t idesc = kk < p.n_main_steps ? idesc_main : idesc_extra;
#pragma unroll
          for (int j = 0; j < MT; ++j) {
            const uint64_t adesc = make_sw128_desc(a_addr + j * A_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 16 elements = 32 bytes along K inside the swizzle row: +2 in the (addr >> 4) field
              umma_bf16(d_tmem + (uint32_t)(j * BN), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, ((kk - kbeg) | k) != 0);
            }
          }
          // slot is free once these MMAs have read it (told to every CTA whose producer writes into it)
          if (CL == 1) umma_commit(empty_bar(stage));
          else umma_commit_mc(empty_bar(stage), (uint16_t)((1u << CL) - 1));
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));  // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      pdl_launch_dependents();  // only the last epilogue is left: the next kernel may be scheduled (it waits for this grid)
    }
  } else {
    // ===================================== epilogue ==========================================
    const int sub = warp & 3;  // TMEM sub-partition this warp may read: lanes 32*sub .. 32*sub+31
    const int row = sub * 32 + lane;
    float* cs_tr = reinterpret_cast<float*>(smem_raw + p.cs_off);  // [4 warps][32][33] transpose scratch
    float* cs_acc = cs_tr + 4 * 32 * 33;                           // [4 warps][B][Cout][2], one slot per warp
    if (p.chsum)
      for (int i = lane; i < p.B * p.Cout * 2; i += 32) cs_acc[(size_t)sub * p.B * p.Cout * 2 + i] = 0.f;
    int acc = 0;
    uint32_t acc_phase = 0;
    WorkIter it(p, CL, crank);
    int tile, kbeg, kend;
    while (it.next(tile, kbeg, kend)) {
      const bool part = kbeg != 0 || kend != p.nk;  // stream-K: this CTA owns only a piece of the tile's K range
      int slot = 0;
      if (part) slot = (int)blockIdx.x - sk_owner((int64_t)tile * p.nk, (int64_t)p.num_tiles * p.nk, (int)gridDim.x);
      const int nt = tile % p.nNt;
      int m = tile / p.nNt;
      const int wt = m % p.nWt; m /= p.nWt;
      const int ht = m % p.nHt; m /= p.nHt;
      const int zt = m % p.nZt;
      const int b = m / p.nZt;
      const int n0 = nt * BN;
      // row -> voxel of the brick (same enumeration as the TMA box: w fastest, then h, then z)
      const int rw = row % p.bw, rh = (row / p.bw) % p.bh, rz = row / (p.bw * p.bh);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
      const int w = wt * p.bw * (p.pw ? MT : 1) + mt * p.pw + rw, h = ht * p.bh * (p.ph ? MT : 1) + mt * p.ph + rh,
                z = zt * p.bz * (p.pz ? MT : 1) + mt * p.pz + rz;
      const bool valid = tile < p.num_tiles && rz < p.bz && w < p.Wo && h < p.Ho && z < p.Z;
      const int64_t vox = (((int64_t)b * p.Z + z) * p.Ho + h) * p.Wo + w;
      const uint32_t t_row = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * ACC_COLS + mt * BN);
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t r[32];
        tmem_ld32(t_row + (uint32_t)c, r);
        // issue the residual / bias loads before waiting for the TMEM load so their latencies overlap
        uint4 rres[4];
        const bool res1 = valid && !part && (p.res_mode == RES_SAME || p.res_mode == RES_UP);
        if (res1) {
          int64_t rrow = vox;
          if (p.res_mode == RES_UP) rrow = (((int64_t)b * p.Z + z) * (p.Ho / 2) + h / 2) * (p.Wo / 2) + w / 2;
          const uint4* rp = reinterpret_cast<const uint4*>((const TS*)p.res + rrow * p.Cout + n0 + c);
#pragma unroll
          for (int j = 0; j < 4; ++j) rres[j] = rp[j];
        }
        tmem_ld_wait();
        if (part) {  // raw fp32 partial sums; bias / residual / rounding happen in the fix-up kernel
          {
            float4* pp = reinterpret_cast<float4*>(
                p.partial + ((((size_t)tile * p.max_slots + slot) * MT + mt) * 128 + row) * BN + c);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              pp[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                  __uint_as_float(r[4 * j + 3]));
          }
          continue;
        }
        float v[32], bv[32];
        if (valid) {
          const float4* bp = reinterpret_cast<const float4*>(p.bias + n0 + c);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = __ldg(bp + j);
            bv[4 * j] = b4.x; bv[4 * j + 1] = b4.y; bv[4 * j + 2] = b4.z; bv[4 * j + 3] = b4.w;
            v[4 * j] = __uint_as_float(r[4 * j]) + b4.x;
            v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b4.y;
            v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b4.z;
            v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b4.w;
          }
          if (res1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) add8r<TS>(v + 8 * j, rres[j]);
          } else if (p.res_mode == RES_POOL) {  // residual = AvgPool(1,2,2) of a (2Ho, 2Wo) tensor
            const int Hr = 2 * p.Ho, Wr = 2 * p.Wo;
            const int64_t r0 = (((int64_t)b * p.Z + z) * Hr + 2 * h) * Wr + 2 * w;
            float s[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) s[j] = 0.f;
            const int64_t offs[4] = {r0, r0 + 1, r0 + Wr, r0 + Wr + 1};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const TS* rp = (const TS*)p.res + offs[q] * p.Cout + n0 + c;
#pragma unroll
              for (int j = 0; j < 4; ++j) add8<TS>(s + 8 * j, rp + 8 * j, 1.0f);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += 0.25f * s[j];
          }
          TS* op = (TS*)p.out + vox * p.Cout + n0 + c;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t w4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) w4[q] = pack2<TS>(v[8 * j + 2 * q], v[8 * j + 2 * q + 1]);
            *reinterpret_cast<uint4*>(op + 8 * j) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) { v[j] = 0.f; bv[j] = 0.f; }
        }
        if (p.chsum) {
          // column sums over the warp's 32 rows through a padded smem transpose (conflict-free both ways):
          // lane = row writes its 32 values, lane = channel reads its column and adds in row order.
          // The sums are taken over (x - bias) (see the strip kernel's epilogue)
          float* tr = cs_tr + sub * (32 * 33);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 32; ++j) tr[lane * 33 + j] = v[j] - bv[j];
          __syncwarp();
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
          for (int rr = 0; rr < 32; rr += 2) {
            const float x0 = tr[rr * 33 + lane], x1 = tr[(rr + 1) * 33 + lane];
            s0 += x0; q0 = fmaf(x0, x0, q0);
            s1 += x1; q1 = fmaf(x1, x1, q1);
          }
          float* acc = cs_acc + (((size_t)sub * p.B + (b < p.B ? b : 0)) * p.Cout + n0 + c + lane) * 2;
          acc[0] += s0 + s1;
          acc[1] += q0 + q1;
        }
      }
      }  // MT
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.chsum) {  // combine the four warps' slots in a fixed order and publish this CTA's partial
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int et = (warp - 2) * 32 + lane, n = p.B * p.Cout * 2;
      for (int i = et; i < n; i += 128) {
        const float t = ((cs_acc[i] + cs_acc[n + i]) + cs_acc[2 * n + i]) + cs_acc[3 * n + i];
        const int bb = i / (p.Cout * 2), rem = i - bb * p.Cout * 2;
        p.chsum[((size_t)bb * p.cs_slots + blockIdx.x) * p.Cout * 2 + rem] = t;
      }
    }
  }

  // ---- teardown ---------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // the peer may still multicast into this CTA's smem / arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// =================================================================================================
// Strip variant for the large 3x3x3 layers: the A operand is staged ONCE per (dz, 64-channel chunk) and reused
// by the 9 in-plane taps.
//
// A z-plane band of width Wb is viewed as a padded, flattened sequence q = h * Wp + w' (Wp = Wb + 2, w' = 0 and
// Wp-1 are the conv's zero padding / the neighbouring band's voxels).  A CTA tile is NB*128 consecutive q.  One
// 5-D TMA box {64 ch, Wp, nh, 1, 1} starting at (w0 - 1, h_s) lands in smem as nh*Wp rows of 128 bytes -- exactly
// that padded-flattened order -- and tap (dh, dw) of brick j is simply the 128 rows starting
// (q0 + 128 j - h_s Wp) + dh Wp + dw rows into the strip.  tcgen05 applies the 128-byte swizzle to absolute smem
// address bits, so an operand descriptor may start at any 128-byte row (probe_rowshift_kernel verifies this on the
// hardware).  Per 9 taps the SM now pulls one 75 KB strip + 9 x 16 KB weight tiles through L2 instead of 9 x 48 KB.
// The ~2/Wp pad columns are computed and discarded.
constexpr int STRIP_THREADS = 224;  // warp 0 strip producer, 1 MMA, 2-5 epilogue, 6 weight producer
constexpr int STRIP_MAXW = 8;

struct StripParams {
  int B, Z, H, W, Cout;
  int Wb, Wp, nbands, nh, tiles_per_band, nNt, num_tiles, zoff;
  int NV;            // voxels (padded-flattened positions) per tile = the MMA N: a multiple of 16, <= 256
  int NP, nZt;       // z-planes per tile (1, or 2 on small planes: both planes share every weight tile) and ceil(Z / NP).
                     // With NP = 2 the two plane accumulators fill the 512 TMEM columns: no double buffering
  int NW;            // stages of the weight ring (3 .. min(STRIP_MAXW, ConvArgs::strip_maxw)) that fit beside the two strips.
                     // Measured: 8 stages instead of 4 on the 48^2 / 24^2 layers change nothing (profiles/r4_ab_options.txt)
  int nsrc, chunks[3], n_macro_main, n_macro, Cin;
  int dz0;           // z offset of the first tap plane: -1 for 3x3x3, 0 for the 3x3 kernels of a dims = 2 network
  int taps;          // 27 or 9
  uint32_t strip_bytes, strip_stride;  // bytes delivered per strip / smem distance between the two strip buffers
  const float* bias;
  const void* res;   // residual (applied after the transpose, voxel-major) or NULL
  int res_mode;
  void* out;
  float* chsum;      // per-CTA channel sums of the output (with a residual: taken after it was added, second transpose)
  int cs_slots;
  uint32_t cs_off;
};

template <typename T, typename TS, int NB, int NP>
__global__ void __launch_bounds__(STRIP_THREADS, 1)
conv_tc_strip_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                     const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapW, const StripParams p) {
  constexpr int BN = 128;
  constexpr int W_BYTES = BN * BK * 2;
  constexpr int ACC_COLS = NB * BN;
  constexpr int TMEM_COLS = 2 * ACC_COLS;
  static_assert(TMEM_COLS == 512 || TMEM_COLS == 256, "TMEM allocation must be a power of two <= 512");
  extern __shared__ uint8_t smem_raw[];
  constexpr int NWB = STRIP_MAXW;  // barrier slots of the weight ring (p.NW of them in use)
  __shared__ __align__(8) uint64_t bars[2 * 2 + 2 * NWB + 4];
  __shared__ uint32_t tmem_base_slot;
  const int NW = p.NW;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = smem_base + 2 * p.strip_stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto sfull = [&](int i) { return bar0 + 8u * i; };
  auto sempty = [&](int i) { return bar0 + 8u * (2 + i); };
  auto wfull = [&](int i) { return bar0 + 8u * (4 + i); };
  auto wempty = [&](int i) { return bar0 + 8u * (4 + NWB + i); };
  auto tfull = [&](int i) { return bar0 + 8u * (4 + 2 * NWB + i); };
  auto tempty = [&](int i) { return bar0 + 8u * (6 + 2 * NWB + i); };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA0);
    prefetch_tmap(&mapW);
    if (p.nsrc > 1) prefetch_tmap(&mapA1);
    if (p.nsrc > 2) prefetch_tmap(&mapA2);
    for (int i = 0; i < 2; ++i) { mbar_init(sfull(i), 1); mbar_init(sempty(i), 1); mbar_init(tfull(i), 1); mbar_init(tempty(i), 4); }
    for (int i = 0; i < NW; ++i) { mbar_init(wfull(i), 1); mbar_init(wempty(i), 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();  // see conv_tc_kernel

  // tile -> (b, z, band, q0, n0)
  auto decode = [&](int tile, int& b, int& z, int& w0, int& q0, int& n0) {
    const int nt = tile % p.nNt;
    int m = tile / p.nNt;
    const int tq = m % p.tiles_per_band; m /= p.tiles_per_band;
    const int band = m % p.nbands; m /= p.nbands;
    z = (m % p.nZt) * NP;  // first plane of the tile
    b = m / p.nZt;
    w0 = band * p.Wb;
    q0 = tq * p.NV;
    n0 = nt * BN;
  };
  // first strip row (may be negative: rows above the plane arrive as zeros)
  auto strip_h0 = [&](int q0) {
    const int x = q0 - p.Wp - 1 + 4 * p.Wp;  // shift into the non-negative range for the division
    return x / p.Wp - 4;
  };

  if (warp == 0) {
    // ===================================== strip producer =====================================
    if (lane == 0) {
      int sb = 0;
      uint32_t sph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int b, z, w0, q0, n0;
        decode(tile, b, z, w0, q0, n0);
        const int h_s = strip_h0(q0);
        for (int m = 0; m < p.n_macro; ++m) {
          mbar_wait(sempty(sb), sph ^ 1);
          mbar_expect_tx(sfull(sb), p.strip_bytes);
          const uint32_t dst = smem_base + sb * p.strip_stride;
          if (m < p.n_macro_main) {
            const int dz = m / p.chunks[0], ch = m - dz * p.chunks[0];
            tma_load_5d(dst, &mapA0, sfull(sb), ch * BK, w0 - 1, h_s, z + dz + p.dz0 + p.zoff, b);
          } else {
            const int e = m - p.n_macro_main;
            if (e < p.chunks[1]) tma_load_5d(dst, &mapA1, sfull(sb), e * BK, w0 - 1, h_s, z, b);
            else tma_load_5d(dst, &mapA2, sfull(sb), (e - p.chunks[1]) * BK, w0 - 1, h_s, z, b);
          }
          if (++sb == 2) { sb = 0; sph ^= 1; }
        }
      }
    }
  } else if (warp == 6) {
    // ===================================== weight producer ====================================
    if (lane == 0) {
      int ws = 0;
      uint32_t wph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n0 = (tile % p.nNt) * BN;
        for (int m = 0; m < p.n_macro; ++m) {
          int ntaps, kcol;
          if (m < p.n_macro_main) {
            const int dz = m / p.chunks[0], ch = m - dz * p.chunks[0];
            ntaps = 9;
            kcol = dz * 9 * p.Cin + ch * BK;  // + t * Cin per in-plane tap
          } else {
            ntaps = 1;
            kcol = p.taps * p.Cin + (m - p.n_macro_main) * BK;
          }
          for (int t = 0; t < ntaps; ++t) {
            mbar_wait(wempty(ws), wph ^ 1);
            mbar_expect_tx(wfull(ws), W_BYTES);
            tma_load_2d(w_base + ws * W_BYTES, &mapW, wfull(ws), kcol + t * p.Cin, n0);
            if (++ws == NW) { ws = 0; wph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =========================================
    if (lane == 0) {
      const uint32_t idesc_main = make_idesc(128, p.NV, !std::is_same<T, f16>::value);
      const uint32_t idesc_extra = make_idesc(128, p.NV, !std::is_same<TS, f16>::value);
      int sb = 0, ws = 0, acc = 0;
      uint32_t sph = 0, wph = 0, aph = 0;
      constexpr int nacc = NP == 2 ? 1 : 2;               // accumulator buffers (of NP * ACC_COLS columns)
      const uint32_t plane_rows = (uint32_t)(p.nh * p.Wp);  // the planes of a strip follow each other (TMA box order)
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int b, z, w0, q0, n0;
        decode(tile, b, z, w0, q0, n0);
        const int offbase = q0 - strip_h0(q0) * p.Wp;  // strip row of the tile's first output position (centre tap)
        mbar_wait(tempty(acc), aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
        bool first = true;
        for (int m = 0; m < p.n_macro; ++m) {
          mbar_wait(sfull(sb), sph);
          tc_fence_after();
          const uint32_t strip = smem_base + sb * p.strip_stride;
          const int ntaps = m < p.n_macro_main ? 9 : 1;
          const uint32_t idesc = m < p.n_macro_main ? idesc_main : idesc_extra;
          for (int t = 0; t < ntaps; ++t) {
            const int dh = ntaps == 9 ? t / 3 - 1 : 0, dw = ntaps == 9 ? t % 3 - 1 : 0;
            mbar_wait(wfull(ws), wph);
            tc_fence_after();
            // Operands are swapped with respect to the brick kernel: M = 128 output channels (the weight tile),
            // N = 256 voxels (256 consecutive strip rows).  One N=256 instruction reads 4 KB + 8 KB of smem per
            // 128 tensor cycles instead of 2 x (4 + 4) KB -- shared-memory read bandwidth (128 B/clk, shared with the
            // TMA fills) is what holds the M=128, N=128 form at ~65 % tensor-pipe utilisation.
            const uint64_t wdesc = make_sw128_desc(w_base + ws * W_BYTES);
#pragma unroll
            for (int pl = 0; pl < NP; ++pl) {  // the planes of the tile share this weight tile
              const uint64_t sdesc = make_sw128_desc(strip + ((uint32_t)pl * plane_rows + (uint32_t)(offbase + dh * p.Wp + dw)) * 128u);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)
                umma_bf16(d_tmem + (uint32_t)(pl * ACC_COLS), wdesc + (uint64_t)(2 * k), sdesc + (uint64_t)(2 * k), idesc,
                          (first && k == 0) ? 0u : 1u);
            }
            first = false;
            umma_commit(wempty(ws));
            if (++ws == NW) { ws = 0; wph ^= 1; }
          }
          umma_commit(sempty(sb));
          if (++sb == 2) { sb = 0; sph ^= 1; }
        }
        umma_commit(tfull(acc));
        if (++acc == nacc) { acc = 0; aph ^= 1; }
      }
      pdl_launch_dependents();
    }
  } else {
    // ===================================== epilogue ===========================================
    // D is [lane = output channel][column = voxel of the tile].  A warp owns 32 channels (its TMEM sub-partition);
    // per 32-voxel chunk it adds bias, accumulates the GroupNorm channel sums (a thread owns a channel: no
    // cross-lane reduction), transposes through padded smem and stores 32 channels x 2 bytes per voxel.
    const int sub = warp & 3;
    float* cs_tr = reinterpret_cast<float*>(smem_raw + p.cs_off);  // [2][32][33]: the warps of a pair take turns
    float* cs_acc = cs_tr + 2 * 32 * 33;                           // [B][Cout][2]
    float* tr = cs_tr + (sub >> 1) * (32 * 33);
    if (p.chsum) {
      for (int i = (warp - 2) * 32 + lane; i < p.B * p.Cout * 2; i += 128) cs_acc[i] = 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    int acc = 0;
    uint32_t aph = 0;
    constexpr int nacc = NP == 2 ? 1 : 2;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int b, z0, w0, q0, n0;
      decode(tile, b, z0, w0, q0, n0);
      const int co = n0 + sub * 32 + lane;
      const float bias_co = __ldg(p.bias + co);
      const bool with_res = p.res_mode != RES_NONE;
      float ts = 0.f, tq = 0.f;
      mbar_wait(tfull(acc), aph);
      tc_fence_after();
#pragma unroll 1
      for (int pl = 0; pl < NP; ++pl) {
      const int z = z0 + pl;
      if (NP > 1 && z >= p.Z) break;  // odd Z: the second plane of the last pair does not exist (uniform over the CTA)
      const uint32_t t_row = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * ACC_COLS + pl * ACC_COLS);
#pragma unroll 1
      for (int vc = 0; vc * 32 < p.NV; ++vc) {
        uint32_t r[32];
        tmem_ld32(t_row + (uint32_t)(vc * 32), r);
        // the voxel this thread will store after the transpose
        const int q = q0 + vc * 32 + lane;
        const int hq = q / p.Wp, wq = q - hq * p.Wp;
        const int w = w0 + wq - 1;
        const bool valid = vc * 32 + lane < p.NV && wq >= 1 && wq <= p.Wb && hq < p.H && w < p.W;
        const int64_t vox = (((int64_t)b * p.Z + z) * p.H + hq) * p.W + w;
        const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
        tmem_ld_wait();
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          // channel sums are taken over (x - bias) = the raw accumulator: a large bias (|mean| >> std) does not cancel
          // in the fp32 sums; the GroupNorm finalize folds the bias back in fp64
          const float a = ((vmask >> j) & 1u) ? __uint_as_float(r[j]) : 0.f;
          x[j] = ((vmask >> j) & 1u) ? a + bias_co : 0.f;
          if (!with_res) {
            ts += a;
            tq = fmaf(a, a, tq);
          }
        }
        float v[32];
#pragma unroll
        for (int turn = 0; turn < 2; ++turn) {
          if ((sub & 1) == turn) {
#pragma unroll
            for (int j = 0; j < 32; ++j) tr[j * 33 + lane] = x[j];  // [voxel][channel]
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] = tr[lane * 33 + c];
          }
          asm volatile("bar.sync %0, 64;" ::"r"(2 + (sub >> 1)) : "memory");
        }
        if (valid) {
          const int cb = n0 + sub * 32;
          if (p.res_mode == RES_SAME) {
            const TS* rp = (const TS*)p.res + vox * p.Cout + cb;
#pragma unroll
            for (int i = 0; i < 4; ++i) add8<TS>(v + 8 * i, rp + 8 * i, 1.0f);
          } else if (p.res_mode == RES_POOL) {  // residual = AvgPool(1,2,2) of a (2H, 2W) tensor
            const int Wr = 2 * p.W;
            const int64_t r0 = (((int64_t)b * p.Z + z) * (2 * p.H) + 2 * hq) * Wr + 2 * w;
            float sacc[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) sacc[i] = 0.f;
            const int64_t offs[4] = {r0, r0 + 1, r0 + Wr, r0 + Wr + 1};
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
              const TS* rp = (const TS*)p.res + offs[qd] * p.Cout + cb;
#pragma unroll
              for (int i = 0; i < 4; ++i) add8<TS>(sacc + 8 * i, rp + 8 * i, 1.0f);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += 0.25f * sacc[i];
          } else if (p.res_mode == RES_UP) {  // residual = nearest x2 of a (H/2, W/2) tensor
            const int64_t r0 = (((int64_t)b * p.Z + z) * (p.H / 2) + hq / 2) * (p.W / 2) + w / 2;
            const TS* rp = (const TS*)p.res + r0 * p.Cout + cb;
#pragma unroll
            for (int i = 0; i < 4; ++i) add8<TS>(v + 8 * i, rp + 8 * i, 1.0f);
          }
          TS* op = (TS*)p.out + vox * p.Cout + cb;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint32_t w4[4];
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) w4[qq] = pack2<TS>(v[8 * i + 2 * qq], v[8 * i + 2 * qq + 1]);
            *reinterpret_cast<uint4*>(op + 8 * i) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
          }
        }
        if (p.chsum && p.res_mode != RES_NONE) {
          // the residual was added voxel-major: transpose the final values back so that lane = channel again and take
          // the sums over (x - bias) of what was stored (the un-rounded fp32 values, as in the other epilogues)
#pragma unroll
          for (int turn = 0; turn < 2; ++turn) {
            if ((sub & 1) == turn) {
#pragma unroll
              for (int c = 0; c < 32; ++c) tr[lane * 33 + c] = v[c];  // [voxel][channel]
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float a = ((vmask >> j) & 1u) ? tr[j * 33 + lane] - bias_co : 0.f;
                ts += a;
                tq = fmaf(a, a, tq);
              }
            }
            asm volatile("bar.sync %0, 64;" ::"r"(2 + (sub >> 1)) : "memory");
          }
        }
      }
      }  // planes of the tile
      if (p.chsum) {  // this thread is the only one of the CTA that owns channel `co`
        float* a2 = cs_acc + ((size_t)b * p.Cout + co) * 2;
        a2[0] += ts;
        a2[1] += tq;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(acc));
      if (++acc == nacc) { acc = 0; aph ^= 1; }
    }
    if (p.chsum) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int n = p.B * p.Cout * 2;
      for (int i = (warp - 2) * 32 + lane; i < n; i += 128) {
        const int bb = i / (p.Cout * 2), rem = i - bb * p.Cout * 2;
        p.chsum[((size_t)bb * p.cs_slots + blockIdx.x) * p.Cout * 2 + rem] = cs_acc[i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}


// stream-K fix-up: for every tile whose K range was shared by several CTAs, add their fp32 partials in slot
// order (deterministic), apply bias / residual and store.  One block per tile.
template <typename T, int MT, int BN>
__global__ void __launch_bounds__(256) sk_fixup_kernel(const TcParams p, int G) {
  const int tile = blockIdx.x;
  const int64_t Ttot = (int64_t)p.num_tiles * p.nk;
  const int c_first = sk_owner((int64_t)tile * p.nk, Ttot, G), c_last = sk_owner((int64_t)(tile + 1) * p.nk - 1, Ttot, G);
  const int nslots = c_last - c_first + 1;
  if (nslots <= 1) return;  // the tile was completed by one CTA's normal epilogue
  const int nt = tile % p.nNt;
  int m = tile / p.nNt;
  const int wt = m % p.nWt; m /= p.nWt;
  const int ht = m % p.nHt; m /= p.nHt;
  const int zt = m % p.nZt;
  const int b = m / p.nZt;
  const int n0 = nt * BN;
  constexpr int NV = BN / 8;
  const float* base = p.partial + (size_t)tile * p.max_slots * MT * 128 * BN;
  // blockIdx.y splits the tile's rows so that a few dozen shared tiles still fill the GPU
  const int per = (MT * 128 * NV + (int)gridDim.y - 1) / (int)gridDim.y;
  const int i_end = min(MT * 128 * NV, ((int)blockIdx.y + 1) * per);
  for (int i = (int)blockIdx.y * per + threadIdx.x; i < i_end; i += blockDim.x) {
    const int rowt = i / NV, c = (i - rowt * NV) * 8;
    const int mt = rowt / 128, row = rowt - mt * 128;
    const int rw = row % p.bw, rh = (row / p.bw) % p.bh, rz = row / (p.bw * p.bh);
    const int w = wt * p.bw * (p.pw ? MT : 1) + mt * p.pw + rw, h = ht * p.bh * (p.ph ? MT : 1) + mt * p.ph + rh,
              z = zt * p.bz * (p.pz ? MT : 1) + mt * p.pz + rz;
    if (!(rz < p.bz && w < p.Wo && h < p.Ho && z < p.Z)) continue;
    const int64_t vox = (((int64_t)b * p.Z + z) * p.Ho + h) * p.Wo + w;
    float v[8];
    {
      const float4 b0 = *reinterpret_cast<const float4*>(p.bias + n0 + c), b1 = *reinterpret_cast<const float4*>(p.bias + n0 + c + 4);
      v[0] = b0.x; v[1] = b0.y; v[2] = b0.z; v[3] = b0.w; v[4] = b1.x; v[5] = b1.y; v[6] = b1.z; v[7] = b1.w;
    }
    for (int sidx = 0; sidx < nslots; ++sidx) {
      const float4* pp = reinterpret_cast<const float4*>(base + ((size_t)sidx * MT * 128 + rowt) * BN + c);
      const float4 a0 = pp[0], a1 = pp[1];
      v[0] += a0.x; v[1] += a0.y; v[2] += a0.z; v[3] += a0.w; v[4] += a1.x; v[5] += a1.y; v[6] += a1.z; v[7] += a1.w;
    }
    const T* res = (const T*)p.res;
    if (p.res_mode == RES_SAME) {
      add8<T>(v, res + vox * p.Cout + n0 + c, 1.0f);
    } else if (p.res_mode == RES_POOL) {
      const int Wr = 2 * p.Wo;
      const int64_t r0 = (((int64_t)b * p.Z + z) * (2 * p.Ho) + 2 * h) * Wr + 2 * w;
      float sacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      add8<T>(sacc, res + r0 * p.Cout + n0 + c, 1.0f);
      add8<T>(sacc, res + (r0 + 1) * p.Cout + n0 + c, 1.0f);
      add8<T>(sacc, res + (r0 + Wr) * p.Cout + n0 + c, 1.0f);
      add8<T>(sacc, res + (r0 + Wr + 1) * p.Cout + n0 + c, 1.0f);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += 0.25f * sacc[j];
    } else if (p.res_mode == RES_UP) {
      const int64_t r0 = (((int64_t)b * p.Z + z) * (p.Ho / 2) + h / 2) * (p.Wo / 2) + w / 2;
      add8<T>(v, res + r0 * p.Cout + n0 + c, 1.0f);
    }
    *reinterpret_cast<uint4*>((T*)p.out + vox * p.Cout + n0 + c) =
        make_uint4(pack2<T>(v[0], v[1]), pack2<T>(v[2], v[3]), pack2<T>(v[4], v[5]), pack2<T>(v[6], v[7]));
  }
}

// ---- host side -----------------------------------------------------------------------------------
// brick {bw, bh, bz} with bw*bh*bz <= 128 that wastes the fewest MMA rows
void choose_brick(int Z, int H, int W, int* bw_, int* bh_, int* bz_) {
  double best = -1.0;
  int bbw = 1, bbh = 1, bbz = 1;
  for (int bw = 1; bw <= std::min(W, BM); ++bw) {
    for (int bh = 1; bh <= std::min(H, BM / bw); ++bh) {
      const int bz = std::min(Z, BM / (bw * bh));
      if (bz < 1) continue;
      const double tiles = (double)ceil_div(W, bw) * ceil_div(H, bh) * ceil_div(Z, bz);
      double eff = ((double)W * H * Z) / (tiles * BM);
      // tie-break: prefer wide-in-W bricks (longer runs of adjacent voxels per TMA box row)
      eff += 1e-6 * bw + 1e-9 * bh;
      if (eff > best) { best = eff; bbw = bw; bbh = bh; bbz = bz; }
    }
  }
  *bw_ = bbw; *bh_ = bbh; *bz_ = bbz;
}

template <typename T, typename TS, int MT, int BN, int NSTAGE, int CL>
int launch_cl(const CUtensorMap* maps, const CUtensorMap& mapW, const TcParams& p, cudaStream_t s, bool pdl) {
  constexpr size_t stage_smem = (size_t)NSTAGE * (MT * A_BYTES + BN * BK * 2) + 1024;
  constexpr size_t smem_max = stage_smem + CS_SMEM_MAX;
  static_assert(smem_max <= 227 * 1024, "shared memory budget");
  static uint64_t configured = 0;
  if (first_use_on_device(&configured)) {
    DD_CUDA(cudaFuncSetAttribute(conv_tc_kernel<T, TS, MT, BN, NSTAGE, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
  }
  int grid;
  if (p.sk) {
    grid = sm_count();
  } else {
    const int numM = p.num_tiles / p.nNt;
    const int n_super = (int)ceil_div(numM, CL) * p.nNt;
    grid = std::min(n_super * CL, (sm_count() / CL) * CL);
  }
  TcParams q = p;
  size_t smem = stage_smem;
  if (q.chsum) {
    q.cs_off = (uint32_t)stage_smem;
    smem += CS_TR_BYTES + (size_t)4 * q.B * q.Cout * 2 * sizeof(float);
    q.cs_slots = chsum_slots();
    if (grid < q.cs_slots)  // slots of CTAs that do not exist must read as zero
      DD_CUDA(cudaMemsetAsync(q.chsum, 0, (size_t)q.B * q.cs_slots * q.Cout * 2 * sizeof(float), s));
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  DD_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<T, TS, MT, BN, NSTAGE, CL>, maps[0], maps[1], maps[2], mapW, q));
  if (p.sk) {
    sk_fixup_kernel<TS, MT, BN><<<dim3(p.num_tiles, 8), 256, 0, s>>>(p, grid);
    DD_CUDA(cudaGetLastError());
  }
  return DDPM3D_OK;
}

// stream-K ranges differ per CTA, so those launches cannot share weight tiles (CL = 1)
template <typename T, typename TS, int MT, int BN, int NSTAGE>
int launch(const CUtensorMap* maps, const CUtensorMap& mapW, const TcParams& p, cudaStream_t s, int cl, bool pdl) {
  if (cl == 2) return launch_cl<T, TS, MT, BN, NSTAGE, 2>(maps, mapW, p, s, pdl);
  return launch_cl<T, TS, MT, BN, NSTAGE, 1>(maps, mapW, p, s, pdl);
}
template <typename T, typename TS>
int launch_plan(const CUtensorMap* maps, const CUtensorMap& mapW, const TcParams& p, cudaStream_t s, int cl, int MT, int BN, bool pdl) {
  if (MT == 2) return launch<T, TS, 2, 128, 4>(maps, mapW, p, s, cl, pdl);
  if (BN == 256) return launch<T, TS, 1, 256, 4>(maps, mapW, p, s, cl, pdl);
  if (BN == 128) return launch<T, TS, 1, 128, 6>(maps, mapW, p, s, cl, pdl);
  return launch<T, TS, 1, 64, 8>(maps, mapW, p, s, cl, pdl);
}

// ---- probe (test-only): does a SWIZZLE_128B K-major operand descriptor work when its start address is an
// arbitrary multiple of 128 bytes (a row shift inside the 8-row swizzle atom)?  D = A[shift : shift+128] * I.
__global__ void __launch_bounds__(128) probe_rowshift_kernel(const __grid_constant__ CUtensorMap mapA,
                                                             const __grid_constant__ CUtensorMap mapI, int rows, int shift,
                                                             int mode, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base, i_smem = base + (uint32_t)rows * 128u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_load = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
  if (threadIdx.x == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_load, (uint32_t)(rows * 128 + 64 * 128));
    tma_load_2d(a_smem, &mapA, bar_load, 0, 0);
    tma_load_2d(i_smem, &mapI, bar_load, 0, 0);
    mbar_wait(bar_load, 0);
    tc_fence_after();
    const uint32_t start = a_smem + (uint32_t)shift * 128u;
    uint64_t adesc = make_sw128_desc(start);
    if (mode == 1) adesc |= (uint64_t)((start >> 7) & 7u) << 49;  // matrix base offset
    const uint64_t bdesc = make_sw128_desc(i_smem);
    const uint32_t idesc = make_idesc(128, 64, true);
    for (int k = 0; k < 4; ++k) umma_bf16(tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k != 0);
    umma_commit(bar_mma);
    mbar_wait(bar_mma, 0);
  }
  __syncthreads();
  tc_fence_after();
  for (int c = 0; c < 64; c += 32) {
    uint32_t r[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

}  // namespace

int probe_rowshift(const void* a, int rows, const void* ident, int shift, int mode, float* out, cudaStream_t s) {
  DD_CHECK(rows >= 136 && rows <= 1024 && rows % 8 == 0 && shift >= 0 && shift + 128 <= rows, DDPM3D_ERR_ARG, "probe: bad geometry");
  CUtensorMap mapA, mapI;
  DD_TRY(make_w_map(&mapA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, a, rows > 256 ? 256 : rows, 64, rows > 256 ? 256 : rows));
  DD_CHECK(rows <= 256, DDPM3D_ERR_ARG, "probe: at most 256 rows (one TMA box)");
  DD_TRY(make_w_map(&mapI, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, ident, 64, 64, 64));
  const size_t smem = (size_t)rows * 128 + 64 * 128 + 1024;
  DD_CUDA(cudaFuncSetAttribute(probe_rowshift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_rowshift_kernel<<<1, 128, smem, s>>>(mapA, mapI, rows, shift, mode, out);
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

namespace {

// SM count the planners assume: the device's, or the one a plan query names (host-side tests of the planning logic run
// without a GPU)
thread_local int g_plan_sms = 0;
inline int plan_sms() { return g_plan_sms > 0 ? g_plan_sms : sm_count(); }

struct TcPlan {
  int bw, bh, bz, pw = 0, ph = 0, pz = 0, nW, nH, nZ, MT = 1, BN = 128, S = 1 /* 2 = stream-K */, max_slots = 0;
};

// Tile configuration and split-K factor from a small cost model (cycles per SM).  Bytes pulled through L2 per
// MMA bound this kernel, so 256 accumulator columns per k-step (two bricks sharing a 128-channel weight tile, or
// one brick x 256 channels) run at ~full rate while 128x128 / 128x64 tiles run at ~0.55 / 0.4 of it.  Layers with
// few tiles split K so that every SM gets work; the fp32 partials cost an extra pass that is charged here.
TcPlan make_plan(const ConvArgs& a, int nk) {
  TcPlan best{};
  choose_brick(a.Z, a.Ho, a.Wo, &best.bw, &best.bh, &best.bz);
  const int nW0 = (int)ceil_div(a.Wo, best.bw), nH0 = (int)ceil_div(a.Ho, best.bh), nZ0 = (int)ceil_div(a.Z, best.bz);
  best.nW = nW0; best.nH = nH0; best.nZ = nZ0;
  const int sms = plan_sms();
  const double rows = (double)a.B * a.Z * a.Ho * a.Wo;
  double best_cost = 1e30;
  struct Cfg { int MT, BN; double cyc_per_kstep; };
  const Cfg cfgs[4] = {{2, 128, 512.0}, {1, 256, 512.0}, {1, 128, 256.0 / 0.55}, {1, 64, 128.0 / 0.40}};
  for (const Cfg& c : cfgs) {
    if (a.Cout % c.BN != 0) continue;
    TcPlan t = best;
    t.MT = c.MT; t.BN = c.BN; t.pw = t.ph = t.pz = 0;
    t.nW = nW0; t.nH = nH0; t.nZ = nZ0;
    if (c.MT == 2) {  // pair bricks along the dimension that wastes the least (an even brick count wastes nothing)
      const int n[3] = {nH0, nZ0, nW0};
      int bd = -1;
      double bw_ = 1e9;
      for (int d = 0; d < 3; ++d) {
        if (n[d] < 2) continue;
        const double waste = (double)(2 * ((n[d] + 1) / 2)) / n[d];
        if (waste < bw_ - 1e-9) { bw_ = waste; bd = d; }
      }
      if (bd < 0 || bw_ > 1.13) continue;
      if (bd == 0) { t.ph = t.bh; t.nH = (nH0 + 1) / 2; }
      else if (bd == 1) { t.pz = t.bz; t.nZ = (nZ0 + 1) / 2; }
      else { t.pw = t.bw; t.nW = (nW0 + 1) / 2; }
    }
    const int64_t tiles = (int64_t)a.B * t.nW * t.nH * t.nZ * (a.Cout / c.BN);
    // whole tiles, wave-quantised
    {
      const double cost = (double)ceil_div(tiles, sms) * ((double)nk * c.cyc_per_kstep + 7000.0);
      if (cost < best_cost) { best_cost = cost; best = t; best.S = 1; best.max_slots = 0; }
    }
    // stream-K: an even share of all k-steps per CTA + the fix-up pass over the shared tiles (~2 per CTA)
    const int64_t T = tiles * nk;
    if (a.splitk_allowed && tiles < 4 * (int64_t)sms && T / sms >= 24) {
      const double share = (double)T / sms;
      const double tile_bytes = (double)c.MT * 128 * c.BN * 4;
      const double fix = 5000.0 + 2.0 * sms * tile_bytes * 2.0 / 2200.0;  // written + read, ~4 TB/s at 1.8 GHz
      const double cost = share * c.cyc_per_kstep + 9000.0 + fix;
      if (cost < best_cost) {
        best_cost = cost; best = t; best.S = 2;
        best.max_slots = (int)(nk / (T / sms)) + 2;
      }
    }
  }
  return best;
}

}  // namespace

namespace {

struct StripPlan {
  int Wb, Wp, nbands, nh, tiles_per_band, nNt, num_tiles, NW, NV, NP, nZt;
  uint32_t strip_bytes, strip_stride;
  size_t smem;
};

// strip_allowed: 0 never; 1 the large layers only (planes >= 24 wide, tiles of 192..256 positions, >= 2 tiles per SM);
// 2 also small planes (12..23 wide): one tile of 160..176 padded-flattened positions per plane and TWO z-planes per
// tile (NP = 2), whose MMAs share every weight tile and whose accumulators fill the 512 TMEM columns.  For those planes
// the brick kernel is bound by the bytes it pulls through L2 per MMA (one 16 KB A tile per tap, no reuse) and runs
// stream-K with an fp32 fix-up pass.  With ONE plane per tile the strip form loses (a 16 KB weight tile per tap and
// 2 x 176 tensor cycles is more than the L2 fabric delivers: 125 vs 110 us on 384 -> 384 at 12 x 12 x 96); with two it
// wins 8-19 % on the 384-channel 12^2 layers (93 vs 110 us) and keeps the GroupNorm channel sums in the epilogue.
// DDPM3D_STRIP_EFF (percent) overrides the acceptance threshold of level 2 (tuning).
bool strip_plan(const ConvArgs& a, bool want_chsum, StripPlan* out) {
  if (!a.strip_allowed || !is_half_dt(a.dt) || (a.taps != 27 && a.taps != 9) || a.stride_hw != 1 || a.out_planar_f32) return false;
  if (a.Cout % 128 != 0 || a.main.C % BK != 0) return false;
  for (int e = 0; e < a.n_extra; ++e)
    if (a.extra[e].C % BK != 0) return false;
  if (a.residual && a.res_mode == RES_UP && (a.Ho % 2 || a.Wo % 2)) return false;
  const bool small_ok = a.strip_allowed >= 2;
  StripPlan t{};
  t.Wb = 0;
  for (int d = std::min(a.Wo, 96); d >= (small_ok ? 12 : 24); --d)
    if (a.Wo % d == 0) { t.Wb = d; break; }
  if (!t.Wb) return false;
  t.Wp = t.Wb + 2;
  t.nbands = a.Wo / t.Wb;
  // tile length NV (= MMA N): the multiple of 16 in [192, 256] that wastes the least of the padded plane
  double best_eff = 0.0;
  for (int nv = 256; nv >= 192; nv -= 16) {
    const int tpb = (int)ceil_div((int64_t)a.Ho * t.Wp, nv);
    const double eff = (double)a.Ho * t.Wb / ((double)tpb * nv) * (nv >= 224 ? 1.0 : 0.97);
    if (eff > best_eff + 0.015) { best_eff = eff; t.NV = nv; t.tiles_per_band = tpb; }  // prefer the longest tile
  }
  bool classic = best_eff >= 0.88;  // pad columns + ragged last tile
  if (!classic) {
    if (!small_ok) return false;
    // shorter tiles: below ~160 positions the 16 KB weight tile per tap (read once per 2 NV tensor cycles) exceeds what
    // an SM can pull through L2
    for (int nv = 176; nv >= 160; nv -= 16) {
      const int tpb = (int)ceil_div((int64_t)a.Ho * t.Wp, nv);
      const double eff = (double)a.Ho * t.Wb / ((double)tpb * nv) * 0.95;
      if (eff > best_eff + 0.015) { best_eff = eff; t.NV = nv; t.tiles_per_band = tpb; }
    }
  }
  t.nh = 3 + (int)ceil_div(t.NV + 1, t.Wp);
  if (t.nh > 256) return false;
  t.nNt = a.Cout / 128;
  t.NP = 1;
  t.nZt = a.Z;
  int64_t tiles = (int64_t)a.B * a.Z * t.nbands * t.tiles_per_band * t.nNt;
  if (tiles >= ((int64_t)1 << 31)) return false;
  const int sms = plan_sms();
  if (tiles < 2 * (int64_t)sms) classic = false;
  if (!classic) {
    if (!small_ok) return false;
    // small planes: two z-planes per tile share every weight tile (with one plane of < 192 positions per tile the
    // 16 KB weight tile per tap is more than the L2 fabric delivers); their accumulators fill the 512 TMEM columns
    t.NP = 2;
    t.nZt = (int)ceil_div(a.Z, 2);
    tiles = (int64_t)a.B * t.nZt * t.nbands * t.tiles_per_band * t.nNt;
    // one wave only: the single accumulator buffer cannot overlap a tile's epilogue with the next tile's MMAs, and over
    // several waves the stream-K brick kernel is faster again (512 -> 512 on 12 x 12 x 160: 396 vs 286 us; the 640-plane
    // volume of the c4 leg on one GPU: +2.7 % on the step)
    if (2 * tiles < sms || tiles > sms) return false;
    // useful positions x how full the last wave is: accept from 0.60 (a 12 x 12 plane with 3 channel tiles: 0.78 x 0.97)
    static const double thr = [] {
      const char* e = getenv("DDPM3D_STRIP_EFF");
      return e ? atof(e) / 100.0 : 0.60;
    }();
    const double wave = (double)tiles / ((double)ceil_div(tiles, sms) * sms) * ((double)a.Z / (2.0 * t.nZt));
    if (best_eff * wave < thr) return false;
  }
  t.num_tiles = (int)tiles;
  t.strip_bytes = (uint32_t)(t.NP * t.nh * t.Wp * 128);
  t.strip_stride = (t.strip_bytes + 1023u) & ~1023u;
  const size_t cs = CS_TR_BYTES / 2 + (want_chsum ? (size_t)a.B * a.Cout * 2 * sizeof(float) : 0);  // the transpose scratch is always needed
  for (int nw = std::max(3, std::min(STRIP_MAXW, a.strip_maxw)); nw >= 3; --nw) {
    t.NW = nw;
    t.smem = (size_t)2 * t.strip_stride + (size_t)nw * 128 * BK * 2 + 1024 + cs;
    if (t.smem + 256 <= 227 * 1024) { *out = t; return true; }
  }
  return false;
}

template <typename T, typename TS>
int launch_strip(const CUtensorMap* maps, const CUtensorMap& mapW, const StripParams& p, const StripPlan& plan, cudaStream_t s,
                 bool pdl) {
  static uint64_t configured = 0;
  if (first_use_on_device(&configured)) {
    DD_CUDA(cudaFuncSetAttribute(conv_tc_strip_kernel<T, TS, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256));
    DD_CUDA(cudaFuncSetAttribute(conv_tc_strip_kernel<T, TS, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256));
  }
  const int grid = std::min(p.num_tiles, sm_count());
  if (p.chsum && grid < p.cs_slots)
    DD_CUDA(cudaMemsetAsync(p.chsum, 0, (size_t)p.B * p.cs_slots * p.Cout * 2 * sizeof(float), s));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(STRIP_THREADS);
  cfg.dynamicSmemBytes = plan.smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (plan.NP == 2) DD_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_strip_kernel<T, TS, 2, 2>, maps[0], maps[1], maps[2], mapW, p));
  else DD_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_strip_kernel<T, TS, 2, 1>, maps[0], maps[1], maps[2], mapW, p));
  return DDPM3D_OK;
}

int conv_tc_strip(ConvArgs& a, const StripPlan& plan, bool chsum, cudaStream_t s) {
  StripParams p{};
  p.B = a.B; p.Z = a.Z; p.H = a.Ho; p.W = a.Wo; p.Cout = a.Cout;
  p.Wb = plan.Wb; p.Wp = plan.Wp; p.nbands = plan.nbands; p.nh = plan.nh; p.tiles_per_band = plan.tiles_per_band;
  p.nNt = plan.nNt; p.num_tiles = plan.num_tiles; p.zoff = a.in_zpad;
  p.NV = plan.NV;
  p.NP = plan.NP;
  p.nZt = plan.nZt;
  p.NW = plan.NW;
  p.nsrc = 1 + a.n_extra;
  p.chunks[0] = a.main.C / BK;
  p.Cin = a.main.C;
  p.taps = a.taps;
  p.dz0 = a.taps == 27 ? -1 : 0;
  p.n_macro_main = (a.taps == 27 ? 3 : 1) * p.chunks[0];
  p.n_macro = p.n_macro_main;
  int Ktot = a.taps * a.main.C;
  for (int e = 0; e < a.n_extra; ++e) {
    p.chunks[1 + e] = a.extra[e].C / BK;
    p.n_macro += p.chunks[1 + e];
    Ktot += a.extra[e].C;
  }
  p.strip_bytes = plan.strip_bytes;
  p.strip_stride = plan.strip_stride;
  p.bias = a.bias;
  p.res = a.residual;
  p.res_mode = a.residual ? a.res_mode : RES_NONE;
  p.out = a.out;
  // with a pooled residual the epilogue (four residual rows per voxel) already takes as long as the tile's MMAs: the second
  // transpose the sums then need made that layer 29 % slower, more than the statistics pass it saves (r4h layer table)
  if (a.residual && a.res_mode == RES_POOL) chsum = false;
  p.chsum = chsum ? a.chsum_out : nullptr;
  p.cs_slots = chsum_slots();
  p.cs_off = (uint32_t)((size_t)2 * plan.strip_stride + (size_t)plan.NW * 128 * BK * 2 + 1024);
  a.chsum_written = chsum ? 1 : 0;
  const CUtensorMapDataType tdt = tmap_dtype(a.dt), tdt_io = tmap_dtype(a.io_dt());
  CUtensorMap maps[3];
  DD_TRY(make_act_map(&maps[0], tdt, a.main.ptr, a.B, a.Z + 2 * a.in_zpad, a.Ho, a.Wo, a.main.C, plan.Wp, plan.nh, plan.NP));
  maps[1] = maps[0];
  maps[2] = maps[0];
  for (int e = 0; e < a.n_extra; ++e)
    DD_TRY(make_act_map(&maps[1 + e], tdt_io, a.extra[e].ptr, a.B, a.Z, a.Ho, a.Wo, a.extra[e].C, plan.Wp, plan.nh, plan.NP));
  CUtensorMap mapW;
  DD_TRY(make_w_map(&mapW, tdt, a.w, a.Cout, a.w_ld ? a.w_ld : Ktot, 128));
  const bool pdl = a.pdl != 0;
  if (a.dt == DDPM3D_BF16 && a.io_dt() == DDPM3D_BF16) return launch_strip<bf16, bf16>(maps, mapW, p, plan, s, pdl);
  if (a.dt == DDPM3D_BF16) return launch_strip<bf16, f16>(maps, mapW, p, plan, s, pdl);
  return launch_strip<f16, f16>(maps, mapW, p, plan, s, pdl);
}

}  // namespace

// Which kernel and tiling conv_tc() would pick for this layer on a device with `sms` SMs (0 = the current device).
// out[8] = {kind: 0 not eligible, 1 brick kernel, 2 brick kernel with stream-K, 3 strip kernel; MT | NP (z-planes per
// strip tile); BN | NV (positions per strip tile = MMA N); tiles; grid; weight-ring stages | split-K slots; k-steps |
// macro steps; strip box rows}.  Pure host arithmetic: nothing is launched and no device is needed when sms > 0.
int conv_tc_plan_query(const ConvArgs& a0, int sms, int* out) {
  ConvArgs a = a0;
  for (int i = 0; i < 8; ++i) out[i] = 0;
  if (!conv_tc_eligible(a)) return DDPM3D_OK;
  g_plan_sms = sms;
  StripPlan sp;
  if (strip_plan(a, true, &sp) || strip_plan(a, false, &sp)) {  // (as conv_tc() does when the engine asks for channel sums)
    const int n = plan_sms();
    g_plan_sms = 0;
    const int macro = (a.taps == 27 ? 3 : 1) * (a.main.C / BK) + (a.n_extra > 0 ? a.extra[0].C / BK : 0) + (a.n_extra > 1 ? a.extra[1].C / BK : 0);
    const int v[8] = {3, sp.NP, sp.NV, sp.num_tiles, std::min(sp.num_tiles, n), sp.NW, macro, sp.nh};
    for (int i = 0; i < 8; ++i) out[i] = v[i];
    return DDPM3D_OK;
  }
  int nk = a.taps * (a.main.C / BK);
  for (int e = 0; e < a.n_extra; ++e) nk += a.extra[e].C / BK;
  const TcPlan plan = make_plan(a, nk);
  const int n = plan_sms();
  g_plan_sms = 0;
  const int tiles = a.B * plan.nW * plan.nH * plan.nZ * (a.Cout / plan.BN);
  const int v[8] = {plan.S > 1 ? 2 : 1, plan.MT, plan.BN, tiles, plan.S > 1 ? n : std::min(tiles, n), plan.max_slots, nk, 0};
  for (int i = 0; i < 8; ++i) out[i] = v[i];
  return DDPM3D_OK;
}

size_t conv_tc_scratch_bytes(const ConvArgs& a0) {
  ConvArgs a = a0;
  a.splitk_allowed = 1;
  if (!conv_tc_eligible(a)) return 0;
  {
    StripPlan sp;
    if (strip_plan(a, false, &sp)) return 0;
  }
  int nk = a.taps * (a.main.C / BK);
  for (int e = 0; e < a.n_extra; ++e) nk += a.extra[e].C / BK;
  const TcPlan plan = make_plan(a, nk);
  if (plan.S == 1) return 0;
  const size_t tiles = (size_t)a.B * plan.nW * plan.nH * plan.nZ * (a.Cout / plan.BN);
  return tiles * plan.max_slots * plan.MT * 128 * plan.BN * sizeof(float);
}

bool conv_tc_eligible(const ConvArgs& a) {
  if (!is_half_dt(a.dt) || !is_half_dt(a.io_dt()) || a.out_planar_f32) return false;
  if (a.stride_hw != 1 && !(a.stride_hw == 2 && a.taps != 1 && a.in_zpad == 0)) return false;
  if (a.dt == DDPM3D_FP16 && a.io_dt() != DDPM3D_FP16) return false;  // built: bf16/bf16, bf16/fp16, fp16/fp16
  if (a.taps != 27 && a.taps != 9 && a.taps != 1) return false;
  if (a.main.C % BK != 0 || a.Cout % 64 != 0) return false;
  for (int e = 0; e < a.n_extra; ++e)
    if (a.extra[e].C % BK != 0) return false;
  auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!aligned(a.main.ptr) || !aligned(a.w) || !aligned(a.out) || !aligned(a.bias)) return false;
  if (a.residual && !aligned(a.residual)) return false;
  if (a.residual && a.res_mode == RES_UP && (a.Ho % 2 || a.Wo % 2)) return false;
  if ((int64_t)a.B * a.Z * a.Ho * a.Wo >= (int64_t)1 << 31) return false;
  return true;
}

int conv_tc(ConvArgs& a, cudaStream_t s) {
  DD_CHECK(conv_tc_eligible(a), DDPM3D_ERR_ARG, "conv_tc: shape not eligible");
  {
    const bool want_cs = a.chsum_out && CS_TR_BYTES + (size_t)4 * a.B * a.Cout * 2 * sizeof(float) <= (size_t)CS_SMEM_MAX;
    StripPlan sp;
    if (strip_plan(a, want_cs, &sp)) return conv_tc_strip(a, sp, want_cs, s);
    if (want_cs && strip_plan(a, false, &sp)) return conv_tc_strip(a, sp, false, s);
  }
  TcParams p{};
  a.chsum_written = 0;
  if (a.chsum_out && CS_TR_BYTES + (size_t)4 * a.B * a.Cout * 2 * sizeof(float) <= (size_t)CS_SMEM_MAX) {
    p.chsum = a.chsum_out;
    a.chsum_written = 1;
  }
  p.nsrc = 1 + a.n_extra;
  p.chunks[0] = a.main.C / BK;
  p.taps = a.taps;
  p.zoff = a.in_zpad;
  p.stride = a.stride_hw;
  p.n_main_steps = a.taps * p.chunks[0];
  p.nk = p.n_main_steps;
  int Ktot = a.taps * a.main.C;
  for (int e = 0; e < a.n_extra; ++e) {
    p.chunks[1 + e] = a.extra[e].C / BK;
    p.nk += p.chunks[1 + e];
    Ktot += a.extra[e].C;
  }
  TcPlan plan = make_plan(a, p.nk);
  p.bw = plan.bw; p.bh = plan.bh; p.bz = plan.bz;
  p.pw = plan.pw; p.ph = plan.ph; p.pz = plan.pz;
  p.nWt = plan.nW; p.nHt = plan.nH; p.nZt = plan.nZ;
  const int BN = plan.BN, MT = plan.MT;
  p.nNt = a.Cout / BN;
  p.num_tiles = a.B * p.nZt * p.nHt * p.nWt * p.nNt;
  p.sk = plan.S > 1;
  p.max_slots = plan.max_slots;
  if (p.sk) {
    DD_CHECK(a.splitk_scratch != nullptr &&
                 a.splitk_bytes >= (size_t)p.num_tiles * plan.max_slots * MT * 128 * BN * sizeof(float),
             DDPM3D_ERR_STATE, "conv_tc: stream-K scratch missing");
    p.partial = a.splitk_scratch;
    p.chsum = nullptr;  // GroupNorm falls back to its own statistics pass for split-K layers
    a.chsum_written = 0;
  }
  p.B = a.B; p.Z = a.Z; p.Ho = a.Ho; p.Wo = a.Wo; p.Cout = a.Cout;
  p.a_tx_bytes = (uint32_t)(p.bw * p.bh * p.bz * BK * 2);
  p.bias = a.bias;
  p.res = a.residual;
  p.res_mode = a.residual ? a.res_mode : RES_NONE;
  p.out = a.out;

  const CUtensorMapDataType tdt = tmap_dtype(a.dt), tdt_io = tmap_dtype(a.io_dt());
  CUtensorMap maps[3];
  DD_TRY(make_act_map(&maps[0], tdt, a.main.ptr, a.B, a.Z + 2 * a.in_zpad, a.Ho * a.stride_hw, a.Wo * a.stride_hw, a.main.C, p.bw,
                      p.bh, p.bz, a.stride_hw));
  maps[1] = maps[0];
  maps[2] = maps[0];
  for (int e = 0; e < a.n_extra; ++e)
    DD_TRY(make_act_map(&maps[1 + e], tdt_io, a.extra[e].ptr, a.B, a.Z, a.Ho, a.Wo, a.extra[e].C, p.bw, p.bh, p.bz));
  CUtensorMap mapW;
  const int cl = (!p.sk && a.cluster_allowed && p.num_tiles / p.nNt >= 2 * sm_count()) ? 2 : 1;
  DD_TRY(make_w_map(&mapW, tdt, a.w, a.Cout, a.w_ld ? a.w_ld : Ktot, BN / cl));
  const bool pdl = a.pdl != 0;
  if (a.dt == DDPM3D_BF16 && a.io_dt() == DDPM3D_BF16) return launch_plan<bf16, bf16>(maps, mapW, p, s, cl, MT, BN, pdl);
  if (a.dt == DDPM3D_BF16) return launch_plan<bf16, f16>(maps, mapW, p, s, cl, MT, BN, pdl);
  return launch_plan<f16, f16>(maps, mapW, p, s, cl, MT, BN, pdl);
}

}  // namespace ddpm3d
