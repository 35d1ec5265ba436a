// The two thin convolutions at the ends of the network:
//
//  * stem:  Conv3d(2 -> C, 3x3x3)   (unet.py:809-811)  K = 54, bound by writing the (B,Z,H,W,C) result.
//           16-bit modes: one tcgen05 tile per 128 voxels, im2col rows built by hand (stem_tc_kernel);
//           fp32 mode: CUDA cores with an smem-staged halo brick (stem_conv_kernel).
//  * head:  Conv3d(C -> 1|2, 3x3x3) (unet.py:993-997)  N = 1|2: fp32 in the reference (it runs after
//           h.type(x.dtype), unet.py:1043-1044), so it stays fp32 FMA on the CUDA cores; bound by smem reads.
//
// The CUDA-core kernels stage a haloed input brick in shared memory (zero padding applied while staging) so every
// input voxel is fetched from L2/HBM once per CTA and reused by the 27 taps from smem.
#include <type_traits>

#include "kernels.h"
#include "tc_ptx.cuh"

namespace ddpm3d {

namespace {

// =================================================================================================
// head: C -> COUT (1 or 2), fp32 in, fp32 planar (NCDHW) out
// =================================================================================================
constexpr int HW_ = 32, HH_ = 8, HZ_ = 4;       // output brick per CTA: 32 x 8 x 4 voxels
constexpr int HCC = 8;                          // channels staged per pass
constexpr int HIW = HW_ + 2, HIH = HH_ + 2, HIZ = HZ_ + 2;
constexpr int HEAD_THREADS = 256;               // lane -> w; warp -> (h half, z)

template <int COUT>
__global__ void __launch_bounds__(HEAD_THREADS, 2)
head_conv_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                 float* __restrict__ out, int B, int Z, int H, int W, int C, int zp) {
  extern __shared__ float4 sm4[];
  float4* s_in = sm4;                                              // [2][HIZ][HIH][HIW] float4
  float* s_w = reinterpret_cast<float*>(sm4 + 2 * HIZ * HIH * HIW);  // [27][COUT][HCC]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nWt = (W + HW_ - 1) / HW_, nHt = (H + HH_ - 1) / HH_, nZt = (Z + HZ_ - 1) / HZ_;
  int t = blockIdx.x;
  const int wt = t % nWt; t /= nWt;
  const int ht = t % nHt; t /= nHt;
  const int zt = t % nZt;
  const int b = t / nZt;
  const int w0 = wt * HW_, h0 = ht * HH_, z0 = zt * HZ_;
  const int lz = warp >> 1, lh0 = (warp & 1) * 4;  // this thread: voxels (lz, lh0..lh0+3, lane)

  float acc[4][COUT];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[i][c] = 0.f;

  const int Ktot = 27 * C;
  for (int c0 = 0; c0 < C; c0 += HCC) {
    __syncthreads();
    // stage the haloed brick: 8 channels (two float4) per voxel
    for (int i = tid; i < HIZ * HIH * HIW * 2; i += HEAD_THREADS) {
      const int q = i & 1;
      int v = i >> 1;
      const int iw = v % HIW; v /= HIW;
      const int ih = v % HIH;
      const int iz = v / HIH;
      const int gz = z0 + iz - 1, gh = h0 + ih - 1, gw = w0 + iw - 1;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gz >= -zp && gz < Z + zp && gh >= 0 && gh < H && gw >= 0 && gw < W)
        val = __ldg(reinterpret_cast<const float4*>(in + ((((int64_t)b * (Z + 2 * zp) + gz + zp) * H + gh) * W + gw) * C + c0) + q);
      s_in[((q * HIZ + iz) * HIH + ih) * HIW + iw] = val;
    }
    for (int i = tid; i < 27 * COUT * HCC; i += HEAD_THREADS) {
      const int c = i % HCC;
      const int co = (i / HCC) % COUT;
      const int tap = i / (HCC * COUT);
      s_w[i] = w[(int64_t)co * Ktot + tap * C + c0 + c];
    }
    __syncthreads();
#pragma unroll 1
    for (int dz = 0; dz < 3; ++dz) {
#pragma unroll
      for (int dw = 0; dw < 3; ++dw) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float4 x[6];
#pragma unroll
          for (int r = 0; r < 6; ++r) x[r] = s_in[((q * HIZ + lz + dz) * HIH + lh0 + r) * HIW + lane + dw];
#pragma unroll
          for (int dh = 0; dh < 3; ++dh) {
            const int tap = (dz * 3 + dh) * 3 + dw;
#pragma unroll
            for (int co = 0; co < COUT; ++co) {
              const float4 wv = *reinterpret_cast<const float4*>(s_w + (tap * COUT + co) * HCC + q * 4);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 xv = x[i + dh];
                acc[i][co] = fmaf(xv.x, wv.x, acc[i][co]);
                acc[i][co] = fmaf(xv.y, wv.y, acc[i][co]);
                acc[i][co] = fmaf(xv.z, wv.z, acc[i][co]);
                acc[i][co] = fmaf(xv.w, wv.w, acc[i][co]);
              }
            }
          }
        }
      }
    }
  }
  const int gz = z0 + lz, gw = w0 + lane;
  if (gz < Z && gw < W) {
    const int64_t sp = (int64_t)Z * H * W;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gh = h0 + lh0 + i;
      if (gh >= H) continue;
      const int64_t pos = ((int64_t)gz * H + gh) * W + gw;
#pragma unroll
      for (int co = 0; co < COUT; ++co) out[((int64_t)b * COUT + co) * sp + pos] = acc[i][co] + bias[co];
    }
  }
}

// =================================================================================================
// head, second version: 32 x 16 x 4 output brick (halo overhead 1.79x instead of 1.99x of L2 traffic), a thread =
// 8 voxels along H x COUT outputs (10 staged float4 feed 192 FMA per (dz, dw) instead of 6 per 96: the first
// version is bound by shared-memory reads), 4 channels per stage, and the stages double-buffered with cp.async
// (zero fill = the conv's padding) so staging overlaps the FMAs instead of alternating with them.
// =================================================================================================
constexpr int VW = 32, VH = 16, VZ = 4;
constexpr int VIW = VW + 2, VIH = VH + 2, VIZ = VZ + 2;
constexpr int VVOX = VIW * VIH * VIZ;                 // 3672 staged voxels
constexpr int V2_THREADS = 256;                       // lane -> w; warp -> (z, h half)
constexpr int VSLOTS = (VVOX + V2_THREADS - 1) / V2_THREADS;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int COUT>
__global__ void __launch_bounds__(V2_THREADS, 1)
head_conv_v2_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                    float* __restrict__ out, int B, int Z, int H, int W, int C, int zp) {
  extern __shared__ float4 sm4[];
  float4* s_in = sm4;                      // [2][VVOX]
  float4* s_w = sm4 + 2 * VVOX;            // [2][27][COUT]
  const uint32_t s_in_u = smem_u32(s_in), s_w_u = smem_u32(s_w);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nWt = (W + VW - 1) / VW, nHt = (H + VH - 1) / VH, nZt = (Z + VZ - 1) / VZ;
  int t = blockIdx.x;
  const int wt = t % nWt; t /= nWt;
  const int ht = t % nHt; t /= nHt;
  const int zt = t % nZt;
  const int b = t / nZt;
  const int w0 = wt * VW, h0 = ht * VH, z0 = zt * VZ;
  const int lz = warp >> 1, lh0 = (warp & 1) * 8;  // this thread: voxels (lz, lh0..lh0+7, lane)

  // the staged voxels of this thread are the same in every channel pass: resolve them once
  int64_t goff[VSLOTS];
#pragma unroll
  for (int k = 0; k < VSLOTS; ++k) {
    const int i = tid + k * V2_THREADS;
    int v = i;
    const int iw = v % VIW; v /= VIW;
    const int ih = v % VIH;
    const int iz = v / VIH;
    const int gz = z0 + iz - 1, gh = h0 + ih - 1, gw = w0 + iw - 1;
    const bool ok = i < VVOX && gz >= -zp && gz < Z + zp && gh >= 0 && gh < H && gw >= 0 && gw < W;
    goff[k] = ok ? ((((int64_t)b * (Z + 2 * zp) + gz + zp) * H + gh) * W + gw) * C : -1;
  }
  const int Ktot = 27 * C;
  auto stage = [&](int c0, int buf) {
#pragma unroll
    for (int k = 0; k < VSLOTS; ++k) {
      const int i = tid + k * V2_THREADS;
      if (i < VVOX) {
        const bool ok = goff[k] >= 0;
        cp_async16(s_in_u + (uint32_t)(buf * VVOX + i) * 16u, ok ? in + goff[k] + c0 : in, ok ? 16 : 0);
      }
    }
    for (int i = tid; i < 27 * COUT; i += V2_THREADS) {
      const int co = i % COUT, tap = i / COUT;
      cp_async16(s_w_u + (uint32_t)(buf * 27 * COUT + i) * 16u, w + (int64_t)co * Ktot + tap * C + c0, 16);
    }
    cp_async_commit();
  };

  float acc[8][COUT];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[i][c] = 0.f;

  const int P = C / 4;
  stage(0, 0);
  for (int p = 0; p < P; ++p) {
    const int buf = p & 1;
    if (p + 1 < P) {
      stage((p + 1) * 4, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float4* xin = s_in + buf * VVOX;
    const float4* wv4 = s_w + buf * 27 * COUT;
#pragma unroll 1
    for (int dz = 0; dz < 3; ++dz) {
#pragma unroll
      for (int dw = 0; dw < 3; ++dw) {
        float4 x[10];
#pragma unroll
        for (int r = 0; r < 10; ++r) x[r] = xin[((lz + dz) * VIH + lh0 + r) * VIW + lane + dw];
#pragma unroll
        for (int dh = 0; dh < 3; ++dh) {
          const int tap = (dz * 3 + dh) * 3 + dw;
#pragma unroll
          for (int co = 0; co < COUT; ++co) {
            const float4 wv = wv4[tap * COUT + co];  // same address for the whole warp: broadcast
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 xv = x[i + dh];
              acc[i][co] = fmaf(xv.x, wv.x, acc[i][co]);
              acc[i][co] = fmaf(xv.y, wv.y, acc[i][co]);
              acc[i][co] = fmaf(xv.z, wv.z, acc[i][co]);
              acc[i][co] = fmaf(xv.w, wv.w, acc[i][co]);
            }
          }
        }
      }
    }
    __syncthreads();  // every warp is done with this stage before it is refilled two passes later
  }
  const int gz = z0 + lz, gw = w0 + lane;
  if (gz < Z && gw < W) {
    const int64_t sp = (int64_t)Z * H * W;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int gh = h0 + lh0 + i;
      if (gh >= H) continue;
      const int64_t pos = ((int64_t)gz * H + gh) * W + gw;
#pragma unroll
      for (int co = 0; co < COUT; ++co) out[((int64_t)b * COUT + co) * sp + pos] = acc[i][co] + bias[co];
    }
  }
}

// =================================================================================================
// stem: 2 -> Cout (multiple of 32), T in / T out, channels-last
// =================================================================================================
constexpr int SW_ = 32, SH_ = 8, SZ_ = 2;       // output brick per CTA: 32 x 8 x 2 voxels = 16 rows of 32
constexpr int SIW = SW_ + 2, SIH = SH_ + 2, SIZ = SZ_ + 2;
constexpr int STEM_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(STEM_THREADS, 2)
stem_conv_kernel(const T* __restrict__ in, const T* __restrict__ w, const float* __restrict__ bias, T* __restrict__ out,
                 int B, int Z, int H, int W, int Cout, int zp) {
  extern __shared__ float4 sm4[];
  float* s_w = reinterpret_cast<float*>(sm4);         // [54][Cout]   (k-major so a channel group is contiguous)
  float* s_b = s_w + 54 * Cout;                       // [Cout]
  float2* s_in = reinterpret_cast<float2*>(s_b + Cout);  // [SIZ][SIH][SIW] (2 channels)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nWt = (W + SW_ - 1) / SW_, nHt = (H + SH_ - 1) / SH_, nZt = (Z + SZ_ - 1) / SZ_;
  int t = blockIdx.x;
  const int wt = t % nWt; t /= nWt;
  const int ht = t % nHt; t /= nHt;
  const int zt = t % nZt;
  const int b = t / nZt;
  const int w0 = wt * SW_, h0 = ht * SH_, z0 = zt * SZ_;

  for (int i = tid; i < 54 * Cout; i += STEM_THREADS) {
    const int n = i % Cout, k = i / Cout;  // packed weight: [Cout][54], k = tap*2 + ci
    s_w[i] = to_f32(w[(int64_t)n * 54 + k]);
  }
  for (int i = tid; i < Cout; i += STEM_THREADS) s_b[i] = bias[i];
  for (int i = tid; i < SIZ * SIH * SIW; i += STEM_THREADS) {
    int v = i;
    const int iw = v % SIW; v /= SIW;
    const int ih = v % SIH;
    const int iz = v / SIH;
    const int gz = z0 + iz - 1, gh = h0 + ih - 1, gw = w0 + iw - 1;
    float2 val = make_float2(0.f, 0.f);
    if (gz >= -zp && gz < Z + zp && gh >= 0 && gh < H && gw >= 0 && gw < W) {
      const T* p = in + ((((int64_t)b * (Z + 2 * zp) + gz + zp) * H + gh) * W + gw) * 2;
      val = make_float2(to_f32(p[0]), to_f32(p[1]));
    }
    s_in[i] = val;
  }
  __syncthreads();

  const int ngroups = Cout / 32;
  const int nwork = SZ_ * SH_ * ngroups;  // (row, channel group) pairs, one per warp pass
  for (int item = warp; item < nwork; item += STEM_THREADS / 32) {
    const int g = item % ngroups;
    const int row = item / ngroups;
    const int lh = row % SH_, lz = row / SH_;
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = s_b[g * 32 + j];
#pragma unroll 1
    for (int tap = 0; tap < 27; ++tap) {
      const int dz = tap / 9, dh = (tap / 3) % 3, dw = tap % 3;
      const float2 v = s_in[((lz + dz) * SIH + lh + dh) * SIW + lane + dw];
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const float xk = ci ? v.y : v.x;
        const float4* wp = reinterpret_cast<const float4*>(s_w + (2 * tap + ci) * Cout + g * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 wv = wp[j];  // same address for the whole warp: broadcast
          acc[4 * j] = fmaf(xk, wv.x, acc[4 * j]);
          acc[4 * j + 1] = fmaf(xk, wv.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(xk, wv.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(xk, wv.w, acc[4 * j + 3]);
        }
      }
    }
    const int gz = z0 + lz, gh = h0 + lh, gw = w0 + lane;
    if (gz < Z && gh < H && gw < W) {
      T* op = out + ((((int64_t)b * Z + gz) * H + gh) * W + gw) * Cout + g * 32;
      constexpr int N = Vec<T>::N;
#pragma unroll
      for (int j = 0; j < 32 / N; ++j) {
        Vec<T> v;
        v.pack(acc + j * N);
        v.store(op + j * N);
      }
    }
  }
}


// =================================================================================================
// stem on the tensor cores (16-bit modes): the 27 taps x 2 channels of a voxel are ONE 128-byte swizzle row
// (K = 54, zero-padded to 64), so a tile of 128 voxels is a single M128 x N(Cout) x K64 tcgen05 accumulation.
// The im2col rows cannot come from TMA (a row gathers 27 scattered 4-byte voxels), so each thread gathers its
// voxel's row from L2 (the packed input is 4 bytes per voxel, L2-resident) and writes it in the SWIZZLE_128B
// layout by hand.  Thread i owns A row i and TMEM lane i.  The kernel is bound by writing the output once;
// several CTAs per SM overlap gather / MMA / store phases.
// =================================================================================================
constexpr int STEM_TC_THREADS = 128;

template <typename T>
__global__ void __launch_bounds__(STEM_TC_THREADS)
stem_tc_kernel(const uint32_t* __restrict__ in, const T* __restrict__ w, const float* __restrict__ bias, T* __restrict__ out,
               int Z, int H, int W, int Cout, int zp, uint32_t tmem_cols, int64_t total, int ntiles, float* __restrict__ chsum,
               int nB) {
  extern __shared__ uint8_t stem_smem_raw[];
  __shared__ uint64_t bar_storage;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[256];
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t raw = smem_u32(stem_smem_raw);
  const uint32_t a_smem = (raw + 1023u) & ~1023u;     // [128 rows][128 B]
  const uint32_t w_smem = a_smem + 128 * 128;         // [Cout rows][128 B]
  uint8_t* a_ptr = stem_smem_raw + (a_smem - raw);
  uint8_t* w_ptr = a_ptr + 128 * 128;
  const uint32_t bar = smem_u32(&bar_storage);
  // GroupNorm channel sums of the output (see ConvArgs::chsum_out): per warp a [32][33] transpose scratch and
  // [nB][Cout][2] accumulators behind the weight rows; the launcher guarantees that a tile never straddles two batch elements
  float* cs_tr = reinterpret_cast<float*>(w_ptr + (size_t)Cout * 128) + warp * (32 * 33);
  float* cs_acc = reinterpret_cast<float*>(w_ptr + (size_t)Cout * 128) + 4 * 32 * 33 + (size_t)warp * nB * Cout * 2;
  if (chsum)
    for (int i = tid & 31; i < nB * Cout * 2; i += 32) cs_acc[i] = 0.f;

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // weights [Cout][54] -> K-major swizzled rows of 64 (k = tap * 2 + ci; 54..63 zero)
  for (int i = tid; i < Cout * 8; i += STEM_TC_THREADS) {
    const int n = i >> 3, c = i & 7;
    uint32_t v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = c * 8 + 2 * q;  // 54 is even: a pair is either fully inside or fully outside
      v[q] = k < 54 ? *reinterpret_cast<const uint32_t*>(w + (int64_t)n * 54 + k) : 0u;
    }
    *reinterpret_cast<uint4*>(w_ptr + n * 128 + ((c ^ (n & 7)) << 4)) = make_uint4(v[0], v[1], v[2], v[3]);
  }
  for (int i = tid; i < Cout; i += STEM_TC_THREADS) s_bias[i] = bias[i];
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = make_idesc(128, Cout, sizeof(T) == 2 && std::is_same<T, bf16>::value);
  const uint64_t adesc = make_sw128_desc(a_smem), wdesc = make_sw128_desc(w_smem);
  const int Zp = Z + 2 * zp;
  uint32_t phase = 0;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t m = (int64_t)tile * 128 + tid;
    const bool live = m < total;
    int x = 0, y = 0, z = 0, b = 0;
    if (live) {
      int64_t r = m;
      x = (int)(r % W); r /= W;
      y = (int)(r % H); r /= H;
      z = (int)(r % Z);
      b = (int)(r / Z);
    }
    uint32_t row[32];
#pragma unroll
    for (int tap = 0; tap < 27; ++tap) {
      const int gz = z + tap / 9 - 1, gy = y + (tap / 3) % 3 - 1, gx = x + tap % 3 - 1;
      const bool ok = live && gz >= -zp && gz < Z + zp && (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
      row[tap] = ok ? __ldg(in + (((int64_t)b * Zp + gz + zp) * H + gy) * W + gx) : 0u;
    }
#pragma unroll
    for (int tap = 27; tap < 32; ++tap) row[tap] = 0u;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      *reinterpret_cast<uint4*>(a_ptr + tid * 128 + ((c ^ (tid & 7)) << 4)) =
          make_uint4(row[4 * c], row[4 * c + 1], row[4 * c + 2], row[4 * c + 3]);
    fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core's async proxy
    tc_fence_before();     // (also orders the previous tile's tcgen05.ld before this tile's MMA)
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < BK / UMMA_K; ++k) umma_bf16(tmem, adesc + 2 * k, wdesc + 2 * k, idesc, k > 0);
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    T* op = out + m * Cout;
    for (int c0 = 0; c0 < Cout; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
      tmem_ld_wait();
      if (live) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t w4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int cc = 8 * j + 2 * q;
            w4[q] = pack2<T>(__uint_as_float(r[cc]) + s_bias[c0 + cc], __uint_as_float(r[cc + 1]) + s_bias[c0 + cc + 1]);
          }
          *reinterpret_cast<uint4*>(op + c0 + 8 * j) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
      }
      if (chsum) {
        // column sums over the warp's 32 voxels of the raw accumulators (= x - bias: a dominant bias does not cancel)
        const int lane = tid & 31;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) cs_tr[lane * 33 + j] = live ? __uint_as_float(r[j]) : 0.f;
        __syncwarp();
        float s0 = 0.f, q0 = 0.f;
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) {
          const float x0 = cs_tr[rr * 33 + lane];
          s0 += x0;
          q0 = fmaf(x0, x0, q0);
        }
        const int tb = (int)(((int64_t)tile * 128) / ((int64_t)Z * H * W));  // batch element of this tile
        float* a2 = cs_acc + ((size_t)tb * Cout + c0 + lane) * 2;
        a2[0] += s0;
        a2[1] += q0;
      }
    }
  }
  if (chsum) {  // combine the four warps' accumulators in a fixed order: this CTA's slot of chsum[nB][gridDim.x][Cout][2]
    __syncthreads();
    const float* acc0 = reinterpret_cast<const float*>(w_ptr + (size_t)Cout * 128) + 4 * 32 * 33;
    const int n = nB * Cout * 2;
    for (int i = tid; i < n; i += STEM_TC_THREADS) {
      const float t = ((acc0[i] + acc0[n + i]) + acc0[2 * n + i]) + acc0[3 * n + i];
      const int bb = i / (Cout * 2), rem = i - bb * Cout * 2;
      chsum[((size_t)bb * gridDim.x + blockIdx.x) * Cout * 2 + rem] = t;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
  }
}

bool stem_tc_eligible(const ConvArgs& a) {
  return is_half_dt(a.dt) && a.stem_tc_allowed && a.Cout % 32 == 0 && a.Cout <= 256 &&
         (reinterpret_cast<uintptr_t>(a.main.ptr) & 3) == 0 && (reinterpret_cast<uintptr_t>(a.w) & 3) == 0;
}

static int stem_tc_grid(const ConvArgs& a, uint32_t* cols_out) {
  const int64_t ntiles = ceil_div((int64_t)a.B * a.Z * a.Ho * a.Wo, 128);
  uint32_t cols = 32;
  while ((int)cols < a.Cout) cols *= 2;
  const int per_sm = std::max(1, std::min(6, 512 / (int)cols));   // TMEM columns bound the co-resident CTAs
  if (cols_out) *cols_out = cols;
  return (int)std::min<int64_t>(ntiles, (int64_t)sm_count() * per_sm);
}

template <typename T>
int stem_tc_launch(ConvArgs& a, cudaStream_t s) {
  const int64_t total = (int64_t)a.B * a.Z * a.Ho * a.Wo;
  const int64_t ntiles = ceil_div(total, 128);
  DD_CHECK(ntiles < ((int64_t)1 << 31), DDPM3D_ERR_ARG, "conv_stem: too many voxels");
  uint32_t cols = 32;
  const int grid = stem_tc_grid(a, &cols);
  // channel sums for the consuming GroupNorm: one slot per CTA (conv_stem_chsum_slots); needs whole tiles per batch element
  float* chsum = conv_stem_chsum_slots(a) ? a.chsum_out : nullptr;
  const size_t smem = 1024 + 128 * 128 + (size_t)a.Cout * 128 + (chsum ? (size_t)4 * 32 * 33 * 4 + (size_t)4 * a.B * a.Cout * 2 * 4 : 0);
  static uint64_t configured = 0;
  if (first_use_on_device(&configured)) {
    DD_CUDA(cudaFuncSetAttribute(stem_tc_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    DD_CUDA(cudaFuncSetAttribute(stem_tc_kernel<f16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  }
  stem_tc_kernel<T><<<grid, STEM_TC_THREADS, smem, s>>>((const uint32_t*)a.main.ptr, (const T*)a.w, a.bias, (T*)a.out, a.Z, a.Ho,
                                                        a.Wo, a.Cout, a.in_zpad, cols, total, (int)ntiles, chsum, a.B);
  a.chsum_written = chsum ? 1 : 0;
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

}  // namespace

bool conv_head_eligible(const ConvArgs& a) {
  return a.out_planar_f32 && a.dt == DDPM3D_FP32 && a.taps == 27 && a.stride_hw == 1 && a.n_extra == 0 && !a.residual &&
         (a.Cout == 1 || a.Cout == 2) && a.main.C % HCC == 0 && (reinterpret_cast<uintptr_t>(a.main.ptr) & 15) == 0;
}

int conv_head(const ConvArgs& a, cudaStream_t s) {
  DD_CHECK(conv_head_eligible(a), DDPM3D_ERR_ARG, "conv_head: shape not eligible");
  if (a.head_v2_allowed && (reinterpret_cast<uintptr_t>(a.w) & 15) == 0) {
    const int grid = a.B * (int)ceil_div(a.Z, VZ) * (int)ceil_div(a.Ho, VH) * (int)ceil_div(a.Wo, VW);
    const size_t smem = (size_t)2 * VVOX * sizeof(float4) + (size_t)2 * 27 * a.Cout * sizeof(float4);
    static uint64_t configured2 = 0;
    if (first_use_on_device(&configured2)) {
      DD_CUDA(cudaFuncSetAttribute(head_conv_v2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
      DD_CUDA(cudaFuncSetAttribute(head_conv_v2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    }
    if (a.Cout == 1)
      head_conv_v2_kernel<1><<<grid, V2_THREADS, smem, s>>>((const float*)a.main.ptr, (const float*)a.w, a.bias, (float*)a.out,
                                                            a.B, a.Z, a.Ho, a.Wo, a.main.C, a.in_zpad);
    else
      head_conv_v2_kernel<2><<<grid, V2_THREADS, smem, s>>>((const float*)a.main.ptr, (const float*)a.w, a.bias, (float*)a.out,
                                                            a.B, a.Z, a.Ho, a.Wo, a.main.C, a.in_zpad);
    DD_CUDA(cudaGetLastError());
    return DDPM3D_OK;
  }
  const int nWt = (int)ceil_div(a.Wo, HW_), nHt = (int)ceil_div(a.Ho, HH_), nZt = (int)ceil_div(a.Z, HZ_);
  const int grid = a.B * nZt * nHt * nWt;
  const size_t smem = (size_t)2 * HIZ * HIH * HIW * sizeof(float4) + (size_t)27 * a.Cout * HCC * sizeof(float);
  static uint64_t configured = 0;
  if (first_use_on_device(&configured)) {
    DD_CUDA(cudaFuncSetAttribute(head_conv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    DD_CUDA(cudaFuncSetAttribute(head_conv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  }
  if (a.Cout == 1)
    head_conv_kernel<1><<<grid, HEAD_THREADS, smem, s>>>((const float*)a.main.ptr, (const float*)a.w, a.bias, (float*)a.out,
                                                         a.B, a.Z, a.Ho, a.Wo, a.main.C, a.in_zpad);
  else
    head_conv_kernel<2><<<grid, HEAD_THREADS, smem, s>>>((const float*)a.main.ptr, (const float*)a.w, a.bias, (float*)a.out,
                                                         a.B, a.Z, a.Ho, a.Wo, a.main.C, a.in_zpad);
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

bool conv_stem_eligible(const ConvArgs& a) {
  return !a.out_planar_f32 && a.main.C == 2 && a.taps == 27 && a.stride_hw == 1 && a.n_extra == 0 && !a.residual &&
         a.Cout % 32 == 0 && a.Cout <= 512 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0;
}

// slots (= CTAs) of the channel-sum buffer the tensor-core stem writes, chsum_out[B][slots][Cout][2]; 0 = this launch
// produces no channel sums (CUDA-core stem, tiles that straddle batch elements, accumulators that do not fit)
int conv_stem_chsum_slots(const ConvArgs& a) {
  if (!conv_stem_eligible(a) || !stem_tc_eligible(a)) return 0;
  if (((int64_t)a.Z * a.Ho * a.Wo) % 128 != 0 || a.B * a.Cout > 512) return 0;
  return stem_tc_grid(a, nullptr);
}

int conv_stem(ConvArgs& a, cudaStream_t s) {
  DD_CHECK(conv_stem_eligible(a), DDPM3D_ERR_ARG, "conv_stem: shape not eligible");
  a.chsum_written = 0;
  if (stem_tc_eligible(a)) return a.dt == DDPM3D_BF16 ? stem_tc_launch<bf16>(a, s) : stem_tc_launch<f16>(a, s);
  const int nWt = (int)ceil_div(a.Wo, SW_), nHt = (int)ceil_div(a.Ho, SH_), nZt = (int)ceil_div(a.Z, SZ_);
  const int grid = a.B * nZt * nHt * nWt;
  const size_t smem = (size_t)(54 + 1) * a.Cout * sizeof(float) + (size_t)SIZ * SIH * SIW * sizeof(float2);
  static uint64_t configured = 0;
  if (first_use_on_device(&configured)) {
    DD_CUDA(cudaFuncSetAttribute(stem_conv_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    DD_CUDA(cudaFuncSetAttribute(stem_conv_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    DD_CUDA(cudaFuncSetAttribute(stem_conv_kernel<f16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  }
  if (a.dt == DDPM3D_BF16)
    stem_conv_kernel<bf16><<<grid, STEM_THREADS, smem, s>>>((const bf16*)a.main.ptr, (const bf16*)a.w, a.bias, (bf16*)a.out,
                                                            a.B, a.Z, a.Ho, a.Wo, a.Cout, a.in_zpad);
  else if (a.dt == DDPM3D_FP16)
    stem_conv_kernel<f16><<<grid, STEM_THREADS, smem, s>>>((const f16*)a.main.ptr, (const f16*)a.w, a.bias, (f16*)a.out,
                                                           a.B, a.Z, a.Ho, a.Wo, a.Cout, a.in_zpad);
  else
    stem_conv_kernel<float><<<grid, STEM_THREADS, smem, s>>>((const float*)a.main.ptr, (const float*)a.w, a.bias,
                                                             (float*)a.out, a.B, a.Z, a.Ho, a.Wo, a.Cout, a.in_zpad);
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

}  // namespace ddpm3d
