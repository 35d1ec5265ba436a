// The network head, fused: out.0 GroupNorm32 apply -> SiLU -> out.2 Conv3d(C -> 1|2, 3x3x3)  (unet.py:993-997, 1043-1044)
// in ONE kernel on the tensor cores, for the 16-bit modes.
//
// The direct form of this convolution is bound by re-reading every activation 27 times from shared memory (N = 2
// outputs give no reuse on the output side): 0.51 ms for 12 GFLOP, plus 0.14 ms for the GroupNorm pass that writes its
// fp32 input.  Here the contraction over channels is done ONCE per voxel and the 27 taps become a shifted sum:
//
//   P[(tap, co)][n] = sum_c w[co][tap][c] * hn[n][c]          one GEMM per z-plane tile: M = 54 (of 128) rows, K = C,
//                                                             N = the (TH+2) x (TW+2) haloed positions of the tile
//   out[co][z][h][w] = sum_tap P[(tap, co)][z + dz - 1][h + dh - 1][w + dw - 1]
//
// * hn = SiLU(x * A[c] + B[c]) (the GroupNorm affine from the finalize kernel) is never stored in HBM: the raw block
//   output x arrives as a TMA box (K-major SWIZZLE_128B rows, zero-filled outside the volume), eight staging warps
//   rewrite it IN PLACE as fp16 hn (zeros at the conv's padding positions) and hand the tile to the MMA warp.
// * P lives in TMEM (double buffered), is copied to shared memory as fp32 and summed with the 27 shifts by the
//   epilogue warps, which keep the three live output planes of the z-sweep in registers.
// * Warp roles: 0-3 epilogue, 4 TMA producer, 5 MMA issuer (+ TMEM allocation), 6-13 staging.
// Operands are fp16 (activations O(1), weights < 1): one rounding of hn and of w to 11 bits, fp32 accumulation -- two
// more roundings of the ~140 a 16-bit evaluation already has.  The fp32 mode keeps the CUDA-core head (conv_small.cu).
#include <type_traits>

#include "kernels.h"
#include "tc_ptx.cuh"

namespace ddpm3d {

namespace {

constexpr int HT_EPI_WARPS = 4, HT_STAGE_WARPS = 8;
constexpr int HT_THREADS = 32 * (HT_EPI_WARPS + 2 + HT_STAGE_WARPS);  // 448
constexpr int HT_NMAX = 256;                    // haloed positions per plane tile (MMA N), <= 256 so TMEM double buffers
constexpr int HT_TILE_BYTES = HT_NMAX * 128;    // one 64-channel chunk of a staged tile
constexpr int HT_W_BYTES = 128 * 128;           // one 64-channel chunk of the stacked weights (128 rows, 54 used)
constexpr int HT_PS = HT_NMAX + 1;              // padded row pitch of the fp32 copy of P (conflict-free both ways)

struct HeadTcParams {
  int B, Z, H, W, C, Cout;
  int TH, TW, PW, NP, N;     // tile, haloed pitch TW+2, haloed positions (TH+2)*(TW+2), N = NP rounded up to 16
  int nHt, nWt, nZr, ZR;     // tiles per plane, z-ranges, planes per z-range
  int num_items;
  int src_f16;               // element format of the block output (1 fp16, 0 bf16)
  const float* ab;           // [B][2][C] GroupNorm affine (A then B)
  const float* w;            // fp32 [Cout][27*C], k = tap*C + c
  const float* bias;         // [Cout]
  float* out;                // planar fp32 (B, Cout, Z, H, W)
};

__global__ void __launch_bounds__(HT_THREADS, 1)
head_tc_kernel(const __grid_constant__ CUtensorMap mapX, const HeadTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[10];
  __shared__ uint32_t tmem_base_slot;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const int chunks = p.C / BK;
  const uint32_t w_smem = base;                                      // [chunks][128 rows][128 B]
  const uint32_t st_smem = w_smem + (uint32_t)chunks * HT_W_BYTES;   // [2 stages][chunks][HT_NMAX rows][128 B]
  const uint32_t stage_bytes = (uint32_t)chunks * HT_TILE_BYTES;
  float* P_s = reinterpret_cast<float*>(base_ptr + (size_t)chunks * HT_W_BYTES + 2 * (size_t)stage_bytes);  // [2][27][HT_PS]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto full = [&](int s) { return bar0 + 8u * s; };          // TMA -> staging
  auto staged = [&](int s) { return bar0 + 8u * (2 + s); };  // staging -> MMA
  auto sfree = [&](int s) { return bar0 + 8u * (4 + s); };   // MMA done with the stage -> TMA
  auto tfull = [&](int a) { return bar0 + 8u * (6 + a); };   // MMA -> epilogue
  auto tfree = [&](int a) { return bar0 + 8u * (8 + a); };   // epilogue -> MMA

  if (threadIdx.x == 0) {
    prefetch_tmap(&mapX);
    for (int i = 0; i < 2; ++i) {
      mbar_init(full(i), 1);
      mbar_init(staged(i), HT_STAGE_WARPS);
      mbar_init(sfree(i), 1);
      mbar_init(tfull(i), 1);
      mbar_init(tfree(i), 2);
    }
    fence_barrier_init();
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // stacked weights: row = co * 32 + tap (rows 27..31, 59..127 zero), K-major SWIZZLE_128B rows of 64 channels, fp16
  for (int i = threadIdx.x; i < chunks * 128 * 8; i += HT_THREADS) {
    const int j = i & 7, row = (i >> 3) & 127, kc = i >> 10;
    const int co = row >> 5, tap = row & 31;
    uint32_t v[4] = {0u, 0u, 0u, 0u};
    if (co < p.Cout && tap < 27) {
      const float* src = p.w + (int64_t)co * 27 * p.C + (int64_t)tap * p.C + kc * BK + j * 8;
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = pack2<f16>(src[2 * q], src[2 * q + 1]);
    }
    *reinterpret_cast<uint4*>(base_ptr + (size_t)kc * HT_W_BYTES + row * 128 + ((j ^ (row & 7)) << 4)) = make_uint4(v[0], v[1], v[2], v[3]);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  auto decode = [&](int item, int& b, int& z_lo, int& z_hi, int& h0, int& w0) {
    const int wt = item % p.nWt; item /= p.nWt;
    const int ht = item % p.nHt; item /= p.nHt;
    const int zr = item % p.nZr;
    b = item / p.nZr;
    z_lo = zr * p.ZR;
    z_hi = min(p.Z, z_lo + p.ZR);
    h0 = ht * p.TH;
    w0 = wt * p.TW;
  };

  if (warp == 4) {
    // ===================================== TMA producer =========================================
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        int b, z_lo, z_hi, h0, w0;
        decode(item, b, z_lo, z_hi, h0, w0);
        const int p_lo = max(0, z_lo - 1), p_hi = min(p.Z, z_hi + 1);
        for (int pz = p_lo; pz < p_hi; ++pz) {
          mbar_wait(sfree(slot), ph ^ 1);
          mbar_expect_tx(full(slot), (uint32_t)(chunks * p.NP * 128));
          for (int kc = 0; kc < chunks; ++kc)
            tma_load_5d(st_smem + slot * stage_bytes + kc * HT_TILE_BYTES, &mapX, full(slot), kc * BK, w0 - 1, h0 - 1, pz, b);
          if (++slot == 2) { slot = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    // ===================================== MMA issuer ===========================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, p.N, false);  // fp16 x fp16 -> fp32
      int slot = 0, acc = 0;
      uint32_t ph = 0, aph = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        int b, z_lo, z_hi, h0, w0;
        decode(item, b, z_lo, z_hi, h0, w0);
        const int p_lo = max(0, z_lo - 1), p_hi = min(p.Z, z_hi + 1);
        for (int pz = p_lo; pz < p_hi; ++pz) {
          mbar_wait(tfree(acc), aph ^ 1);
          mbar_wait(staged(slot), ph);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * HT_NMAX);
          for (int kc = 0; kc < chunks; ++kc) {
            const uint64_t wdesc = make_sw128_desc(w_smem + kc * HT_W_BYTES);
            const uint64_t xdesc = make_sw128_desc(st_smem + slot * stage_bytes + kc * HT_TILE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_bf16(d_tmem, wdesc + (uint64_t)(2 * k), xdesc + (uint64_t)(2 * k), idesc, (kc | k) != 0 ? 1u : 0u);
          }
          umma_commit(sfree(slot));
          umma_commit(tfull(acc));
          if (++slot == 2) { slot = 0; ph ^= 1; }
          if (++acc == 2) { acc = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp >= 6) {
    // ===================================== staging: x -> fp16 SiLU(x A + B), in place =============
    const int sid = threadIdx.x - 6 * 32;       // 0 .. 255
    const int ngroups = p.C / 8;                // 16-byte channel groups per row (8 or 16)
    const int cg = sid % ngroups, rl = sid / ngroups, nrl = (HT_STAGE_WARPS * 32) / ngroups;
    const int kc = cg >> 3, jl = cg & 7;
    int slot = 0;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      int b, z_lo, z_hi, h0, w0;
      decode(item, b, z_lo, z_hi, h0, w0);
      const int p_lo = max(0, z_lo - 1), p_hi = min(p.Z, z_hi + 1);
      float A[8], Bv[8];
      {
        const float* pa = p.ab + (int64_t)b * 2 * p.C + cg * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) { A[k] = pa[k]; Bv[k] = pa[p.C + k]; }
      }
      for (int pz = p_lo; pz < p_hi; ++pz) {
        mbar_wait(full(slot), ph);
        uint8_t* tile = base_ptr + (size_t)chunks * HT_W_BYTES + (size_t)slot * stage_bytes + (size_t)kc * HT_TILE_BYTES;
        int r = rl / p.PW, c = rl - r * p.PW;
        for (int n = rl; n < p.NP; n += nrl) {
          const int gh = h0 - 1 + r, gw = w0 - 1 + c;
          uint4* q = reinterpret_cast<uint4*>(tile + n * 128 + ((jl ^ (n & 7)) << 4));
          uint4 o = make_uint4(0u, 0u, 0u, 0u);  // the convolution pads hn with zeros, not with SiLU(B)
          if ((unsigned)gh < (unsigned)p.H && (unsigned)gw < (unsigned)p.W) {
            const uint4 v = *q;
            const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
            uint32_t ov[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float x0, x1;
              unpack2_rt(wv[i], p.src_f16, x0, x1);
              ov[i] = pack2<f16>(silu_fast(fmaf(x0, A[2 * i], Bv[2 * i])), silu_fast(fmaf(x1, A[2 * i + 1], Bv[2 * i + 1])));
            }
            o = make_uint4(ov[0], ov[1], ov[2], ov[3]);
          }
          *q = o;
          c += nrl;
          while (c >= p.PW) { c -= p.PW; ++r; }
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's reads of the tile
        __syncwarp();
        if (lane == 0) mbar_arrive(staged(slot));
        if (++slot == 2) { slot = 0; ph ^= 1; }
      }
    }
  } else {
    // ===================================== epilogue: P -> shifted sums -> planar fp32 =============
    const int et = warp * 32 + lane;  // 0 .. 127
    const int nout = p.TH * p.TW;     // <= 196
    int acc = 0;
    uint32_t aph = 0;
    const float bias0 = p.bias[0], bias1 = p.Cout > 1 ? p.bias[1] : 0.f;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      int b, z_lo, z_hi, h0, w0;
      decode(item, b, z_lo, z_hi, h0, w0);
      const int p_lo = max(0, z_lo - 1), p_hi = min(p.Z, z_hi + 1);
      // the three live output planes of this thread's (up to two) output voxels: [voxel][plane offset -1, 0, +1][co]
      float a[2][3][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) { a[i][k][0] = 0.f; a[i][k][1] = 0.f; }
      auto emit = [&](int oz, int k) {  // plane slot k of every voxel of this thread -> out[:, oz]
        if (oz < z_lo || oz >= z_hi) return;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int o = et + i * 128;
          if (o >= nout) continue;
          const int oh = o / p.TW, ow = o - oh * p.TW;
          const int gh = h0 + oh, gw = w0 + ow;
          if (gh >= p.H || gw >= p.W) continue;
          const int64_t sp = (int64_t)p.Z * p.H * p.W;
          const int64_t pos = ((int64_t)oz * p.H + gh) * p.W + gw;
          p.out[(int64_t)b * p.Cout * sp + pos] = a[i][k][0] + bias0;
          if (p.Cout > 1) p.out[((int64_t)b * p.Cout + 1) * sp + pos] = a[i][k][1] + bias1;
        }
      };
      for (int pz = p_lo; pz < p_hi; ++pz) {
        mbar_wait(tfull(acc), aph);
        tc_fence_after();
        if (warp < 2) {  // rows co * 32 + tap live in TMEM lanes 0 .. 63: sub-partitions 0 and 1
          const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * HT_NMAX);
          for (int c0 = 0; c0 < p.N; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(t_row + (uint32_t)c0, r);
            tmem_ld_wait();
            if (lane < 27) {
              float* dst = P_s + (size_t)(warp * 27 + lane) * HT_PS + c0;
#pragma unroll
              for (int j = 0; j < 32; ++j) dst[j] = __uint_as_float(r[j]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tfree(acc));
        }
        if (++acc == 2) { acc = 0; aph ^= 1; }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // input plane pz feeds output planes pz+1 (dz = 0), pz (dz = 1) and pz-1 (dz = 2)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int o = et + i * 128;
          if (o < nout) {
            const int oh = o / p.TW, ow = o - oh * p.TW;
            const float* pb = P_s + oh * p.PW + ow;
#pragma unroll
            for (int dz = 0; dz < 3; ++dz) {
              float s0 = 0.f, s1 = 0.f;
#pragma unroll
              for (int dh = 0; dh < 3; ++dh)
#pragma unroll
                for (int dw = 0; dw < 3; ++dw) {
                  const int tap = (dz * 3 + dh) * 3 + dw;
                  s0 += pb[(size_t)tap * HT_PS + dh * p.PW + dw];
                  s1 += pb[(size_t)(27 + tap) * HT_PS + dh * p.PW + dw];
                }
              a[i][2 - dz][0] += s0;
              a[i][2 - dz][1] += s1;
            }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // P_s may be overwritten by the next plane
        emit(pz - 1, 0);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          a[i][0][0] = a[i][1][0]; a[i][0][1] = a[i][1][1];
          a[i][1][0] = a[i][2][0]; a[i][1][1] = a[i][2][1];
          a[i][2][0] = 0.f; a[i][2][1] = 0.f;
        }
      }
      emit(p_hi - 1, 0);  // the last plane of the volume has no input plane after it
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

struct HeadPlan {
  int TH, TW, nZr, ZR;
};

HeadPlan head_plan(int B, int Z, int H, int W) {
  HeadPlan best{1, 1, 1, Z};
  double best_eff = -1.0;
  for (int th = 1; th <= std::min(H, 62); ++th)
    for (int tw = 1; tw <= std::min(W, 62); ++tw) {
      const int np = (th + 2) * (tw + 2);
      if (np > HT_NMAX || th * tw > 256) continue;
      const int n16 = (np + 15) / 16 * 16;
      const double tiles = (double)ceil_div(H, th) * ceil_div(W, tw);
      const double eff = (double)H * W / (tiles * n16) + 1e-6 * tw;
      if (eff > best_eff) { best_eff = eff; best.TH = th; best.TW = tw; }
    }
  const int64_t per_range = (int64_t)B * ceil_div(H, best.TH) * ceil_div(W, best.TW);
  const int sms = sm_count();
  double best_cost = 1e30;
  for (int nzr = 1; nzr <= Z; ++nzr) {
    const int zr = (int)ceil_div(Z, nzr);
    const int nzr_eff = (int)ceil_div(Z, zr);
    const double cost = (double)ceil_div(per_range * nzr_eff, sms) * (zr + 2 + 1.5);  // + pipeline fill per item
    if (cost < best_cost - 1e-9) { best_cost = cost; best.nZr = nzr_eff; best.ZR = zr; }
  }
  return best;
}

}  // namespace

bool conv_head_tc_eligible(int dt_src, int C, int Cout) {
  return is_half_dt(dt_src) && (C == 64 || C == 128) && (Cout == 1 || Cout == 2);
}

// x: the block output [B][Z][H][W][C] (dt_src, 16 bit); ab: [B][2][C] GroupNorm affine; w / bias: out.2 in fp32
int conv_head_tc(int dt_src, const void* x, const float* ab, const float* w, const float* bias, float* out, int B, int Z, int H,
                 int W, int C, int Cout, cudaStream_t s) {
  DD_CHECK(conv_head_tc_eligible(dt_src, C, Cout), DDPM3D_ERR_ARG, "conv_head_tc: shape not eligible");
  const HeadPlan plan = head_plan(B, Z, H, W);
  HeadTcParams p{};
  p.B = B; p.Z = Z; p.H = H; p.W = W; p.C = C; p.Cout = Cout;
  p.TH = plan.TH; p.TW = plan.TW; p.PW = plan.TW + 2;
  p.NP = (plan.TH + 2) * (plan.TW + 2);
  p.N = (p.NP + 15) / 16 * 16;
  p.nHt = (int)ceil_div(H, plan.TH); p.nWt = (int)ceil_div(W, plan.TW);
  p.nZr = plan.nZr; p.ZR = plan.ZR;
  const int64_t items = (int64_t)B * p.nZr * p.nHt * p.nWt;
  DD_CHECK(items < ((int64_t)1 << 31), DDPM3D_ERR_ARG, "conv_head_tc: too many tiles");
  p.num_items = (int)items;
  p.src_f16 = dt_src == DDPM3D_FP16;
  p.ab = ab; p.w = w; p.bias = bias; p.out = out;
  CUtensorMap mapX;
  DD_TRY(make_act_map(&mapX, tmap_dtype(dt_src), x, B, Z, H, W, C, p.PW, plan.TH + 2, 1));
  const int chunks = C / BK;
  const size_t smem = 1024 + (size_t)chunks * HT_W_BYTES + 2 * (size_t)chunks * HT_TILE_BYTES + (size_t)2 * 27 * HT_PS * sizeof(float);
  static uint64_t configured = 0;
  if (first_use_on_device(&configured))
    DD_CUDA(cudaFuncSetAttribute(head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256));
  DD_CHECK(smem <= 227 * 1024 - 256, DDPM3D_ERR_STATE, "conv_head_tc: shared memory budget");
  const int grid = (int)std::min<int64_t>(items, sm_count());
  head_tc_kernel<<<grid, HT_THREADS, smem, s>>>(mapX, p);
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

}  // namespace ddpm3d
