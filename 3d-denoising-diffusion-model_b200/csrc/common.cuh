// Shared helpers for the ddpm3d CUDA library (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>

#include "../../include/ddpm3d.h"

namespace ddpm3d {

void set_error(const std::string& msg);

#define DD_CUDA(expr)                                                                           \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      ::ddpm3d::set_error(std::string(#expr) + ": " + cudaGetErrorString(e__) + " (" __FILE__ + \
                          ":" + std::to_string(__LINE__) + ")");                                \
      return DDPM3D_ERR_CUDA;                                                                   \
    }                                                                                           \
  } while (0)

#define DD_CHECK(cond, code, msg)         \
  do {                                    \
    if (!(cond)) {                        \
      ::ddpm3d::set_error(msg);           \
      return (code);                      \
    }                                     \
  } while (0)

#define DD_TRY(expr)            \
  do {                          \
    int r__ = (expr);           \
    if (r__ != DDPM3D_OK) return r__; \
  } while (0)

typedef __nv_bfloat16 bf16;
typedef __half f16;

__host__ __device__ inline bool is_half_dt(int dt) { return dt == DDPM3D_BF16 || dt == DDPM3D_FP16; }

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// cudaFuncSetAttribute is per device: a launcher keeps one `uint64_t` mask and asks once per device
inline bool first_use_on_device(uint64_t* mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = 1ull << (dev & 63);
  if (*mask & bit) return false;
  *mask |= bit;
  return true;
}

// SM count of the current device (persistent grids, launch sizing, per-CTA statistics slots): queried, never assumed
inline int sm_count() {
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int& c = n[dev & 63];
  if (c == 0) {
    cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev);
    if (c <= 0) c = 1;
  }
  return c;
}

// programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may begin
// before its predecessor in the stream has finished; pdl_wait() blocks until that predecessor has completed and its
// writes are visible.  pdl_launch_dependents() in the predecessor lets the dependent grid be scheduled right away.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- scalar conversions ------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f32(f16 v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ f16 from_f32<f16>(float v) { return __float2half_rn(v); }

// two packed 16-bit elements <-> two floats
template <typename T> __device__ __forceinline__ void unpack2(uint32_t w, float& a, float& b);
template <> __device__ __forceinline__ void unpack2<bf16>(uint32_t w, float& a, float& b) {
  a = __uint_as_float(w << 16);
  b = __uint_as_float(w & 0xffff0000u);
}
template <> __device__ __forceinline__ void unpack2<f16>(uint32_t w, float& a, float& b) {
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
  a = f.x;
  b = f.y;
}
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<bf16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<f16>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---- 16-bit elements whose format (bf16 / fp16) is a run-time property ---------------------------------
// The CUDA-core convolution reads tensors of two 16-bit formats in one launch (bf16 conv operands next to fp16
// block inputs / outputs, see DESIGN.md "storage formats"); both are 2 bytes, only the conversion differs.
struct h16 { uint16_t raw; };
__device__ __forceinline__ float h16_to_f32(uint16_t raw, int is_f16) {
  return is_f16 ? __half2float(__ushort_as_half(raw)) : __uint_as_float((uint32_t)raw << 16);
}
__device__ __forceinline__ uint16_t f32_to_h16(float v, int is_f16) {
  return is_f16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ void unpack2_rt(uint32_t w, int is_f16, float& a, float& b) {
  if (is_f16) unpack2<f16>(w, a, b);
  else unpack2<bf16>(w, a, b);
}
__device__ __forceinline__ uint32_t pack2_rt(float a, float b, int is_f16) { return is_f16 ? pack2<f16>(a, b) : pack2<bf16>(a, b); }

// ---- 16-byte vectors of activations ---------------------------------------------------------
// A "vec" is 16 bytes: 4 floats or 8 bf16.  All channel counts on the path are multiples of 8
// except the 2-channel network input / output, which have their own kernels.
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  float4 raw;
  __device__ __forceinline__ void load(const float* p) { raw = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = raw; }
  __device__ __forceinline__ void unpack(float* f) const { f[0] = raw.x; f[1] = raw.y; f[2] = raw.z; f[3] = raw.w; }
  __device__ __forceinline__ void pack(const float* f) { raw = make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct Vec<bf16> {
  static constexpr int N = 8;
  uint4 raw;
  __device__ __forceinline__ void load(const bf16* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(bf16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ void unpack(float* f) const {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ __forceinline__ void pack(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    raw = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

template <> struct Vec<f16> {
  static constexpr int N = 8;
  uint4 raw;
  __device__ __forceinline__ void load(const f16* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(f16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ void unpack(float* f) const {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) unpack2<f16>(w[i], f[2 * i], f[2 * i + 1]);
  }
  __device__ __forceinline__ void pack(const float* f) {
    raw = make_uint4(pack2<f16>(f[0], f[1]), pack2<f16>(f[2], f[3]), pack2<f16>(f[4], f[5]), pack2<f16>(f[6], f[7]));
  }
};

// accurate expf + IEEE division: fp32 outputs (fp32 mode, the head's input)
__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + expf(-v)); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kNegLog2e = -1.4426950408889634f;
// 16-bit outputs: v * rcp(1 + ex2(-v log2 e)) with the two raw SFU instructions (<= 3 ulp fp32, far below the 2^-9 /
// 2^-12 output rounding; unbiased): FMUL + MUFU + FADD + MUFU + FMUL.  (__expf / __fdividef compile to the same two
// MUFUs plus a range check of three more instructions per element, which made the GroupNorm apply pass issue-bound.)
// The one-SFU form h + h*tanh.approx(h), h = v/2, was measured 8 % faster on that pass but its 2^-11 error is a BIAS
// (same sign for all positive activations) that the next convolution sums coherently, so it is not used.
__device__ __forceinline__ float silu_fast2(float v, float v2) { return v * rcp_approx(1.0f + ex2_approx(v2)); }
__device__ __forceinline__ float silu_fast(float v) { return silu_fast2(v, v * kNegLog2e); }
template <typename T> __device__ __forceinline__ float silu_t(float v) { return sizeof(T) == 2 ? silu_fast(v) : silu_f(v); }

}  // namespace ddpm3d
