// Generic implicit-GEMM 3-D convolution on the CUDA cores (fp32 accumulate).
//
// This is (a) the convolution of the fp32 parity mode (DDPM3D_FP32: true fp32 FMA, needed for the
// <=1e-4 tolerance the north star states for fp32), (b) the kernel for the few shapes the tcgen05
// kernel does not take (2-channel network input, 2-channel fp32 network output, strided
// Downsample conv, channel counts that are not multiples of 64), and (c) the on-device
// cross-check for conv_tc.cu.  The dominant bf16 convolutions run on conv_tc.cu.
//
// GEMM view: M = B*Z*Ho*Wo output voxels, N = Cout, K = taps*Cmain (+ C of up to two 1x1x1
// "extra" sources: the ResBlock skip_connection folded into the same accumulation, reading the
// two halves of the decoder concat in place).  CTA tile 128x64x16, 256 threads, 8x4 per thread.
#include "kernels.h"

namespace ddpm3d {

namespace {

constexpr int BM = 128, BN = 64, BK = 16, THREADS = 256;

struct SimtParams {
  const void* src[3];
  int C[3];
  int kbeg[4];  // K range of each source; kbeg[nsrc] = Ktot
  int nsrc;
  int taps, stride;
  const void* w;
  const float* bias;
  const void* res;
  int res_mode;
  void* out;
  int planar;
  int B, Z, Ho, Wo, Hin, Win, Cout, Ktot, wld;
  int zp;  // halo planes of source 0
  int64_t M;
  // 16-bit launches: element format (1 = fp16, 0 = bf16) of the main source + its weight columns, and of the
  // extra sources + their weight columns + residual + output
  int f16_main, f16_io;
};

template <typename T> struct Ld8;  // 8 consecutive elements -> 8 floats
template <> struct Ld8<float> {
  __device__ static void ld(const float* p, float* f, int) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};
template <> struct Ld8<h16> {
  __device__ static void ld(const h16* p, float* f, int is_f16) {
    const uint4 raw = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) unpack2_rt(w[i], is_f16, f[2 * i], f[2 * i + 1]);
  }
};
template <typename T> struct Ld4;
template <> struct Ld4<float> {
  __device__ static void ld(const float* p, float* f, int) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  }
};
template <> struct Ld4<h16> {
  __device__ static void ld(const h16* p, float* f, int is_f16) {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    unpack2_rt(r.x, is_f16, f[0], f[1]);
    unpack2_rt(r.y, is_f16, f[2], f[3]);
  }
};
__device__ __forceinline__ float ld1(const float* p, int) { return *p; }
__device__ __forceinline__ float ld1(const h16* p, int is_f16) { return h16_to_f32(p->raw, is_f16); }

// VEC: every source C % 8 == 0 and Ktot % 4 == 0 -> a 16-wide K chunk never straddles a tap/source
template <typename T, bool VEC>
__global__ void __launch_bounds__(THREADS) conv_simt_kernel(SimtParams p) {
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // ---- A loader: thread -> (row, 8-wide k half) -------------------------------------------------
  const int a_row = tid % BM, a_kh = tid / BM;
  const int64_t am = m0 + a_row;
  const bool a_ok = am < p.M;
  int ab = 0, az = 0, aho = 0, awo = 0;
  if (a_ok) {
    awo = (int)(am % p.Wo);
    int64_t t = am / p.Wo;
    aho = (int)(t % p.Ho);
    t /= p.Ho;
    az = (int)(t % p.Z);
    ab = (int)(t / p.Z);
  }
  // ---- B loader: thread -> (n, 4-wide k quarter) ------------------------------------------------
  const int b_n = tid / 4, b_kq = tid % 4;
  const bool b_ok = (n0 + b_n) < p.Cout;
  const T* wrow = (const T*)p.w + (int64_t)(n0 + b_n) * p.wld;

  float ra[8], rb[4];

  auto load_tiles = [&](int k0) {
    // A
    if (VEC) {
#pragma unroll
      for (int j = 0; j < 8; ++j) ra[j] = 0.f;
      const int k = k0 + a_kh * 8;
      if (a_ok && k < p.Ktot) {
        int s = 0;
        while (s + 1 < p.nsrc && k >= p.kbeg[s + 1]) ++s;
        const int kk = k - p.kbeg[s];
        const int C = p.C[s];
        int tap, ci;
        if (s == 0 && p.taps > 1) { tap = kk / C; ci = kk - tap * C; if (p.taps == 9) tap += 9; } else { tap = 13; ci = kk; }
        const int dz = tap / 9 - 1, dh = (tap / 3) % 3 - 1, dw = tap % 3 - 1;
        const int st = s == 0 ? p.stride : 1;
        const int Hs = s == 0 ? p.Hin : p.Ho, Ws = s == 0 ? p.Win : p.Wo;
        const int zpl = s == 0 ? p.zp : 0, Zs = p.Z + 2 * zpl;
        const int zi = az + dz + zpl, hi = aho * st + dh, wi = awo * st + dw;
        if (zi >= 0 && zi < Zs && hi >= 0 && hi < Hs && wi >= 0 && wi < Ws) {
          const T* src = (const T*)p.src[s] + ((((int64_t)ab * Zs + zi) * Hs + hi) * Ws + wi) * C + ci;
          Ld8<T>::ld(src, ra, s == 0 ? p.f16_main : p.f16_io);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ra[j] = 0.f;
        const int k = k0 + a_kh * 8 + j;
        if (a_ok && k < p.Ktot) {
          int s = 0;
          while (s + 1 < p.nsrc && k >= p.kbeg[s + 1]) ++s;
          const int kk = k - p.kbeg[s];
          const int C = p.C[s];
          int tap, ci;
          if (s == 0 && p.taps > 1) { tap = kk / C; ci = kk - tap * C; if (p.taps == 9) tap += 9; } else { tap = 13; ci = kk; }
          const int dz = tap / 9 - 1, dh = (tap / 3) % 3 - 1, dw = tap % 3 - 1;
          const int st = s == 0 ? p.stride : 1;
          const int Hs = s == 0 ? p.Hin : p.Ho, Ws = s == 0 ? p.Win : p.Wo;
          const int zpl = s == 0 ? p.zp : 0, Zs = p.Z + 2 * zpl;
          const int zi = az + dz + zpl, hi = aho * st + dh, wi = awo * st + dw;
          if (zi >= 0 && zi < Zs && hi >= 0 && hi < Hs && wi >= 0 && wi < Ws)
            ra[j] = ld1((const T*)p.src[s] + ((((int64_t)ab * Zs + zi) * Hs + hi) * Ws + wi) * C + ci, s == 0 ? p.f16_main : p.f16_io);
        }
      }
    }
    // B
    const int kb = k0 + b_kq * 4;
    if (VEC) {
#pragma unroll
      for (int j = 0; j < 4; ++j) rb[j] = 0.f;
      if (b_ok && kb < p.Ktot) Ld4<T>::ld(wrow + kb, rb, kb < p.kbeg[1] ? p.f16_main : p.f16_io);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) rb[j] = (b_ok && kb + j < p.Ktot) ? ld1(wrow + kb + j, kb + j < p.kbeg[1] ? p.f16_main : p.f16_io) : 0.f;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 8; ++j) As[buf][a_kh * 8 + j][a_row] = ra[j];
#pragma unroll
    for (int j = 0; j < 4; ++j) Bs[buf][b_kq * 4 + j][b_n] = rb[j];
  };

  const int tx = tid % 16, ty = tid / 16;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = (p.Ktot + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue: + bias (+ residual) -> store ------------------------------------------------------
  const int nb = n0 + tx * 4;
  float bias[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) bias[j] = (p.bias && nb + j < p.Cout) ? p.bias[nb + j] : 0.f;
  const bool vec_out = (p.Cout % 4 == 0) && !p.planar;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty * 8 + i;
    if (m >= p.M) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias[j];
    int wo = 0, ho = 0, z = 0, b = 0;
    if (p.res_mode >= RES_POOL || p.planar) {
      wo = (int)(m % p.Wo);
      int64_t t = m / p.Wo;
      ho = (int)(t % p.Ho);
      t /= p.Ho;
      z = (int)(t % p.Z);
      b = (int)(t / p.Z);
    }
    if (p.res_mode != RES_NONE && nb < p.Cout) {
      const T* r = (const T*)p.res;
      if (p.res_mode == RES_SAME) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (nb + j < p.Cout) v[j] += ld1(r + m * p.Cout + nb + j, p.f16_io);
      } else if (p.res_mode == RES_POOL) {
        const int Hr = 2 * p.Ho, Wr = 2 * p.Wo;
        const int64_t r0 = (((int64_t)b * p.Z + z) * Hr + 2 * ho) * Wr + 2 * wo;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (nb + j < p.Cout) {
            const int c = nb + j;
            // avg_pool3d sums the 4 taps then scales; same here
            const float sum = ld1(r + r0 * p.Cout + c, p.f16_io) + ld1(r + (r0 + 1) * p.Cout + c, p.f16_io) +
                              ld1(r + (r0 + Wr) * p.Cout + c, p.f16_io) + ld1(r + (r0 + Wr + 1) * p.Cout + c, p.f16_io);
            v[j] += 0.25f * sum;
          }
      } else {  // RES_UP
        const int Hr = p.Ho / 2, Wr = p.Wo / 2;
        const int64_t r0 = (((int64_t)b * p.Z + z) * Hr + ho / 2) * Wr + wo / 2;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (nb + j < p.Cout) v[j] += ld1(r + r0 * p.Cout + nb + j, p.f16_io);
      }
    }
    if (p.planar) {
      float* o = (float*)p.out;
      const int64_t sp = (int64_t)p.Z * p.Ho * p.Wo;
      const int64_t pos = ((int64_t)z * p.Ho + ho) * p.Wo + wo;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (nb + j < p.Cout) o[((int64_t)b * p.Cout + nb + j) * sp + pos] = v[j];
    } else if (vec_out) {
      if (nb < p.Cout) {
        T* o = (T*)p.out + m * p.Cout + nb;
        if constexpr (sizeof(T) == 4) {
          *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
          uint2 pk;
          pk.x = pack2_rt(v[0], v[1], p.f16_io);
          pk.y = pack2_rt(v[2], v[3], p.f16_io);
          *reinterpret_cast<uint2*>(o) = pk;
        }
      }
    } else {
      T* o = (T*)p.out + m * p.Cout;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (nb + j < p.Cout) {
          if constexpr (sizeof(T) == 4) o[nb + j] = v[j];
          else o[nb + j].raw = f32_to_h16(v[j], p.f16_io);
        }
    }
  }
}

}  // namespace

int conv_simt(const ConvArgs& a, cudaStream_t s) {
  DD_CHECK(a.taps == 27 || a.taps == 9 || a.taps == 1, DDPM3D_ERR_ARG, "conv: taps must be 27, 9 (3x3 in the plane, dims = 2) or 1");
  DD_CHECK(a.stride_hw == 1 || a.stride_hw == 2, DDPM3D_ERR_ARG, "conv: stride_hw must be 1 or 2");
  DD_CHECK(a.n_extra >= 0 && a.n_extra <= 2, DDPM3D_ERR_ARG, "conv: at most two extra sources");
  SimtParams p{};
  p.nsrc = 1 + a.n_extra;
  p.src[0] = a.main.ptr;
  p.C[0] = a.main.C;
  p.kbeg[0] = 0;
  p.kbeg[1] = a.taps * a.main.C;
  bool vec = (a.main.C % 8 == 0);
  for (int e = 0; e < a.n_extra; ++e) {
    p.src[1 + e] = a.extra[e].ptr;
    p.C[1 + e] = a.extra[e].C;
    p.kbeg[2 + e] = p.kbeg[1 + e] + a.extra[e].C;
    vec = vec && (a.extra[e].C % 8 == 0);
  }
  p.Ktot = p.kbeg[p.nsrc];
  p.wld = a.w_ld ? a.w_ld : p.Ktot;
  vec = vec && (p.Ktot % 4 == 0) && (p.wld % 4 == 0);
  p.taps = a.taps;
  p.stride = a.stride_hw;
  p.zp = a.in_zpad;
  p.w = a.w;
  p.bias = a.bias;
  p.res = a.residual;
  p.res_mode = a.residual ? a.res_mode : RES_NONE;
  p.out = a.out;
  p.planar = a.out_planar_f32;
  p.B = a.B; p.Z = a.Z; p.Ho = a.Ho; p.Wo = a.Wo;
  p.Hin = a.Ho * a.stride_hw; p.Win = a.Wo * a.stride_hw;
  p.Cout = a.Cout;
  p.M = (int64_t)a.B * a.Z * a.Ho * a.Wo;
  DD_CHECK(!(p.res_mode == RES_UP) || (a.Ho % 2 == 0 && a.Wo % 2 == 0), DDPM3D_ERR_ARG, "conv: RES_UP needs even output H, W");
  DD_CHECK(!a.out_planar_f32 || a.dt == DDPM3D_FP32, DDPM3D_ERR_ARG, "conv: planar output is fp32 only");
  dim3 grid((unsigned)ceil_div(p.M, BM), (unsigned)ceil_div(a.Cout, BN));
  if (is_half_dt(a.dt)) {
    p.f16_main = a.dt == DDPM3D_FP16;
    p.f16_io = a.io_dt() == DDPM3D_FP16;
    if (vec) conv_simt_kernel<h16, true><<<grid, THREADS, 0, s>>>(p);
    else conv_simt_kernel<h16, false><<<grid, THREADS, 0, s>>>(p);
  } else {
    if (vec) conv_simt_kernel<float, true><<<grid, THREADS, 0, s>>>(p);
    else conv_simt_kernel<float, false><<<grid, THREADS, 0, s>>>(p);
  }
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

}  // namespace ddpm3d
