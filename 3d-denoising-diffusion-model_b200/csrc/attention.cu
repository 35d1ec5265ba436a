// QKV attention core (unet.py:328-393) on channels-last tokens, online softmax in fp32.
//
// Attention is OFF at the shipped flags (attention_resolutions=1000 -> ds=0 never matches and the
// middle-block attention is commented out, unet.py:876-882).  This CUDA-core flash-style kernel serves fp32
// mode and head widths other than 64; 16-bit models with 64-wide heads run attention_tc.cu (tcgen05).
#include "kernels.h"

namespace ddpm3d {

namespace {

constexpr int QB = 128;  // queries per CTA, one per thread
constexpr int KT = 32;   // keys per smem tile

template <typename T, int CH>
__global__ void __launch_bounds__(QB) attention_kernel(const T* __restrict__ qkv, T* __restrict__ out, int T_tok, int C,
                                                       int heads, int ch, int new_order, int q_begin, int Tq) {
  __shared__ float ks[KT][CH];
  __shared__ float vs[KT][CH];
  const int bh = blockIdx.y, b = bh / heads, h = bh % heads;
  const int qoff = new_order ? h * ch : h * 3 * ch;
  const int koff = new_order ? C + h * ch : h * 3 * ch + ch;
  const int voff = new_order ? 2 * C + h * ch : h * 3 * ch + 2 * ch;
  const int64_t row_stride = 3 * (int64_t)C;
  const T* base = qkv + (int64_t)b * T_tok * row_stride;
  const int tq = blockIdx.x * QB + threadIdx.x;  // query index inside the window
  const int t = q_begin + tq;
  const bool live = tq < Tq;
  const float scale = 1.0f / sqrtf(sqrtf((float)ch));

  float q[CH], acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    q[c] = (live && c < ch) ? to_f32(base[(int64_t)t * row_stride + qoff + c]) * scale : 0.f;
    acc[c] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < T_tok; j0 += KT) {
    __syncthreads();
    for (int i = threadIdx.x; i < KT * CH; i += QB) {
      const int j = i / CH, c = i % CH;
      const bool ok = (j0 + j) < T_tok && c < ch;
      ks[j][c] = ok ? to_f32(base[(int64_t)(j0 + j) * row_stride + koff + c]) * scale : 0.f;
      vs[j][c] = ok ? to_f32(base[(int64_t)(j0 + j) * row_stride + voff + c]) : 0.f;
    }
    __syncthreads();
    const int jn = min(KT, T_tok - j0);
    for (int j = 0; j < jn; ++j) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < CH; ++c) s = fmaf(q[c], ks[j][c], s);
      const float mn = fmaxf(m, s);
      const float corr = expf(m - mn), p = expf(s - mn);
      l = l * corr + p;
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] = acc[c] * corr + p * vs[j][c];
      m = mn;
    }
  }
  if (live) {
    const float inv = 1.0f / l;
    T* o = out + ((int64_t)b * Tq + tq) * C + h * ch;
    for (int c = 0; c < ch; ++c) o[c] = from_f32<T>(acc[c] * inv);
  }
}

template <typename T>
int launch(const void* qkv, void* out, int B, int T_tok, int C, int heads, int new_order, int q_begin, int Tq, cudaStream_t s) {
  const int ch = C / heads;
  dim3 grid((unsigned)ceil_div(Tq, QB), (unsigned)(B * heads));
  if (ch <= 16) attention_kernel<T, 16><<<grid, QB, 0, s>>>((const T*)qkv, (T*)out, T_tok, C, heads, ch, new_order, q_begin, Tq);
  else if (ch <= 32) attention_kernel<T, 32><<<grid, QB, 0, s>>>((const T*)qkv, (T*)out, T_tok, C, heads, ch, new_order, q_begin, Tq);
  else if (ch <= 64) attention_kernel<T, 64><<<grid, QB, 0, s>>>((const T*)qkv, (T*)out, T_tok, C, heads, ch, new_order, q_begin, Tq);
  else if (ch <= 128) attention_kernel<T, 128><<<grid, QB, 0, s>>>((const T*)qkv, (T*)out, T_tok, C, heads, ch, new_order, q_begin, Tq);
  else { set_error("attention: head width > 128 unsupported"); return DDPM3D_ERR_ARG; }
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

}  // namespace

int attention_k(int dt, const void* qkv, void* out, int B, int T_tok, int C, int heads, int new_order, void* scratch,
                size_t scratch_bytes, cudaStream_t s, int q_begin, int q_count) {
  DD_CHECK(heads > 0 && C % heads == 0, DDPM3D_ERR_ARG, "attention: C must be divisible by heads");
  if (q_count < 0) { q_begin = 0; q_count = T_tok; }
  DD_CHECK(q_begin >= 0 && q_count >= 1 && q_begin + q_count <= T_tok, DDPM3D_ERR_ARG, "attention: query window out of range");
  const size_t need = attention_tc_scratch_bytes(dt, B, T_tok, C, heads);
  if (need > 0 && scratch && scratch_bytes >= need) return attention_tc(dt, qkv, out, B, T_tok, C, heads, new_order, scratch, s, q_begin, q_count);
  if (dt == DDPM3D_BF16) return launch<bf16>(qkv, out, B, T_tok, C, heads, new_order, q_begin, q_count, s);
  if (dt == DDPM3D_FP16) return launch<f16>(qkv, out, B, T_tok, C, heads, new_order, q_begin, q_count, s);
  return launch<float>(qkv, out, B, T_tok, C, heads, new_order, q_begin, q_count, s);
}

}  // namespace ddpm3d
