// HBM-bound kernels of the path: input packing, GroupNorm32(+FiLM)(+SiLU)(+pool/upsample),
// timestep embedding MLPs and the p_sample posterior update.  sm_100a; fp32 math everywhere.
#include "kernels.h"

namespace ddpm3d {

// =================================================================================================
// pack_input: cat([x, low_res], dim=1) + cast (unet.py:1690-1693, :1035)
// =================================================================================================
template <typename T>
__global__ void pack_input_kernel(const float* __restrict__ x, const float* __restrict__ low, T* __restrict__ out, int64_t n,
                                  int64_t per_b, int64_t pad_vox, T* __restrict__ peer_lo, T* __restrict__ peer_hi) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const int64_t b = pad_vox ? i / per_b : 0;
    const int64_t o = i + (2 * b + 1) * pad_vox;  // skip the leading halo planes of batches 0..b
    const T vx = from_f32<T>(x[i]), vl = from_f32<T>(low[i]);
    out[2 * o] = vx;
    out[2 * o + 1] = vl;
    if (pad_vox) {  // peer path (one halo plane): first / last plane -> the neighbours' trailing / leading halo plane
      const int64_t sp = i - b * per_b, bo = b * (per_b + 2 * pad_vox);
      if (peer_lo && sp < pad_vox) { peer_lo[2 * (bo + sp)] = vx; peer_lo[2 * (bo + sp) + 1] = vl; }
      if (peer_hi && sp >= per_b - pad_vox) {
        const int64_t q = bo + sp - (per_b - pad_vox);
        peer_hi[2 * q] = vx;
        peer_hi[2 * q + 1] = vl;
      }
    }
  }
}

int pack_input(int dt, const float* x, const float* low, void* out, int B, int Z, int64_t plane, int out_zpad, cudaStream_t s,
               void* peer_lo, void* peer_hi) {
  const int64_t n = (int64_t)B * Z * plane, per_b = (int64_t)Z * plane, pad_vox = (int64_t)out_zpad * plane;
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>(ceil_div(n, threads), sm_count() * 16);
  if (dt == DDPM3D_BF16)
    pack_input_kernel<bf16><<<blocks, threads, 0, s>>>(x, low, (bf16*)out, n, per_b, pad_vox, (bf16*)peer_lo, (bf16*)peer_hi);
  else if (dt == DDPM3D_FP16)
    pack_input_kernel<f16><<<blocks, threads, 0, s>>>(x, low, (f16*)out, n, per_b, pad_vox, (f16*)peer_lo, (f16*)peer_hi);
  else
    pack_input_kernel<float><<<blocks, threads, 0, s>>>(x, low, (float*)out, n, per_b, pad_vox, (float*)peer_lo, (float*)peer_hi);
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

// general form: x and (optionally) low_res with Cx channels each, planar (B, Cx, Z, H, W) -> channels-last
// (B, Z, H, W, Cx [+ Cx]); used by the 2-D / multi-channel model classes (unet.py:396-716, 1650-1673)
template <typename T>
__global__ void pack_planar_kernel(const float* __restrict__ x, const float* __restrict__ low, int Cx, T* __restrict__ out,
                                   int64_t n, int64_t per_b, int64_t pad_vox) {
  const int Ctot = low ? 2 * Cx : Cx;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const int64_t b = i / per_b, sp = i - b * per_b;
    const int64_t o = (i + (2 * b + 1) * pad_vox) * Ctot;
    for (int c = 0; c < Cx; ++c) {
      out[o + c] = from_f32<T>(x[(b * Cx + c) * per_b + sp]);
      if (low) out[o + Cx + c] = from_f32<T>(low[(b * Cx + c) * per_b + sp]);
    }
  }
}

int pack_input_planar(int dt, const float* x, const float* low, int Cx, void* out, int B, int Z, int64_t plane, int out_zpad,
                      cudaStream_t s) {
  const int64_t n = (int64_t)B * Z * plane, per_b = (int64_t)Z * plane, pad_vox = (int64_t)out_zpad * plane;
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>(ceil_div(n, threads), sm_count() * 16);
  if (dt == DDPM3D_BF16) pack_planar_kernel<bf16><<<blocks, threads, 0, s>>>(x, low, Cx, (bf16*)out, n, per_b, pad_vox);
  else if (dt == DDPM3D_FP16) pack_planar_kernel<f16><<<blocks, threads, 0, s>>>(x, low, Cx, (f16*)out, n, per_b, pad_vox);
  else pack_planar_kernel<float><<<blocks, threads, 0, s>>>(x, low, Cx, (float*)out, n, per_b, pad_vox);
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

// =================================================================================================
// GroupNorm32 (nn.py:17-19,93-100): statistics in fp32 over (C/32, Z, H, W) per batch element.
// Pass 1 (stats): per-chunk partial [sum, sumsq] per group, fixed summation order (deterministic).
// Pass 2 (finalize): fp64 reduction over chunks -> per-(b,c) affine  y = x*A + B  with gamma/beta,
//                    FiLM (unet.py:248-252) and the additive-embedding variant folded in.
// Pass 3 (apply):    y -> SiLU -> optional AvgPool(1,2,2) / nearest x2 -> store.
// Algorithmic HBM bytes: 2 reads + 1 write of the tensor.
// =================================================================================================
int gn_chunks(int64_t rows) {
  // about four stats CTAs per SM at full size; never fewer than 64 rows per chunk
  int64_t c = rows / 64;
  if (c < 1) c = 1;
  if (c > 592) c = 592;
  return (int)c;
}

// thread -> (row lane r, channel vector v); every thread keeps its channel vector for the whole kernel,
// so a warp always touches whole rows (>= 64 contiguous bytes) and no index arithmetic sits in the loop.
// Cancellation-safe: a thread accumulates sum(x - p) and sum((x - p)^2) in fp32 around a per-thread, per-channel pivot
// p (the first value it reads), so the running sums stay O(std) however large |mean| / std is; the pivot is folded back
// in fp64 (sum x = s + n p, sum x^2 = q + 2 p s + n p^2) and everything downstream (smem reduction, per-chunk partials,
// finalize) is fp64.  GroupNorm of a tensor with mean / std = 1000 is as accurate as with mean 0
// (tests/test_gpu_kernels.py::test_groupnorm_large_mean).
template <typename T>
__global__ void __launch_bounds__(256) gn_stats_kernel(const T* __restrict__ s0, const T* __restrict__ s1, int C0, int C1,
                                                       int rows, int n_chunks, const float* __restrict__ pre_add,
                                                       int64_t pre_stride, double* __restrict__ partials) {
  constexpr int N = Vec<T>::N;
  extern __shared__ double smd[];  // [rpi][Ctot][2]
  const int Ctot = C0 + C1;
  const int nvec0 = C0 / N, nvec = Ctot / N;
  const int rpi = blockDim.x / nvec;
  const int v = threadIdx.x % nvec, r = threadIdx.x / nvec;
  const int b = blockIdx.y, chunk = blockIdx.x;

  const T* base;
  int Csrc;
  if (v < nvec0) { base = s0 + (int64_t)b * rows * C0 + v * N; Csrc = C0; }
  else { base = s1 + (int64_t)b * rows * C1 + (v - nvec0) * N; Csrc = C1; }

  float add[N];
#pragma unroll
  for (int i = 0; i < N; ++i) add[i] = 0.f;
  if (pre_add) {
#pragma unroll
    for (int i = 0; i < N; ++i) add[i] = pre_add[(int64_t)b * pre_stride + v * N + i];
  }

  float s[N], q[N], piv[N];
#pragma unroll
  for (int i = 0; i < N; ++i) { s[i] = 0.f; q[i] = 0.f; piv[i] = 0.f; }
  int n = 0;
  // Row blocks of U*rpi rows are dealt round-robin to the chunks, so at any moment the whole grid streams one
  // narrow window of the tensor (DRAM page / TLB locality on multi-GB tensors); each chunk still sums its rows
  // in a fixed order.
  constexpr int U = 8;  // independent 16-byte loads in flight per thread
  const int RB = U * rpi;
  {
    const int row0 = chunk * RB + r;  // the first row this thread will read (if any): its values are the pivots
    if (row0 < rows) {
      Vec<T> a;
      a.load(base + (int64_t)row0 * Csrc);
      a.unpack(piv);
#pragma unroll
      for (int i = 0; i < N; ++i) piv[i] += add[i];
    }
  }
  for (int rb0 = chunk * RB; rb0 < rows; rb0 += n_chunks * RB) {
    const int row = rb0 + r;
    if (rb0 + RB <= rows) {
      Vec<T> a[U];
#pragma unroll
      for (int u = 0; u < U; ++u) a[u].load(base + (int64_t)(row + u * rpi) * Csrc);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float f[N];
        a[u].unpack(f);
#pragma unroll
        for (int i = 0; i < N; ++i) { const float x = (f[i] + add[i]) - piv[i]; s[i] += x; q[i] = fmaf(x, x, q[i]); }
      }
      n += U;
    } else {
      for (int rr = row; rr < rows; rr += rpi) {
        Vec<T> a;
        a.load(base + (int64_t)rr * Csrc);
        float f[N];
        a.unpack(f);
#pragma unroll
        for (int i = 0; i < N; ++i) { const float x = (f[i] + add[i]) - piv[i]; s[i] += x; q[i] = fmaf(x, x, q[i]); }
        ++n;
      }
    }
  }
  double* mine = smd + ((int64_t)r * Ctot + v * N) * 2;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double p = (double)piv[i], sd = (double)s[i];
    mine[2 * i] = sd + (double)n * p;
    mine[2 * i + 1] = (double)q[i] + 2.0 * p * sd + (double)n * p * p;
  }
  __syncthreads();
  // group g sums its channels over the rpi row-lanes in a fixed order (deterministic)
  const int gpc = Ctot / 32;
  if (threadIdx.x < 64) {
    const int g = threadIdx.x >> 1, which = threadIdx.x & 1;
    double acc = 0.0;
    for (int rr = 0; rr < rpi; ++rr)
      for (int c = g * gpc; c < (g + 1) * gpc; ++c) acc += smd[((int64_t)rr * Ctot + c) * 2 + which];
    partials[(((int64_t)b * 32 + g) * 2 + which) * n_chunks + chunk] = acc;
  }
}

// partials: [B][32 groups][2][n_chunks]
__global__ void __launch_bounds__(1024) gn_finalize_kernel(const double* __restrict__ partials, int n_chunks, int Ctot,
                                                           double inv_count, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, const float* __restrict__ film,
                                                           int64_t film_stride, const float* __restrict__ pre_add,
                                                           int64_t pre_stride, float* __restrict__ ab) {
  __shared__ float s_mean[32], s_rstd[32];
  const int b = blockIdx.x;
  const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;  // 1024 threads: one warp per group
  pdl_launch_dependents();
  const double* ps = partials + (((int64_t)b * 32 + g) * 2) * n_chunks;
  const double* pq = ps + n_chunks;
  double s = 0.0, q = 0.0;
  for (int c = lane; c < n_chunks; c += 32) {  // coalesced; fixed order -> deterministic
    s += ps[c];
    q += pq[c];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane == 0) {
    const double mean = s * inv_count;
    double var = q * inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[g] = (float)mean;
    s_rstd[g] = (float)(1.0 / sqrt(var + 1e-5));
  }
  __syncthreads();
  const int gpc = Ctot / 32;
  float* A = ab + (int64_t)b * 2 * Ctot;
  float* Bv = A + Ctot;
  for (int c = threadIdx.x; c < Ctot; c += blockDim.x) {
    const int gg = c / gpc;
    float a = s_rstd[gg] * gamma[c];
    float o = beta[c] - s_mean[gg] * a;
    if (film) {  // h = norm(h) * (1 + scale) + shift
      const float sc = 1.0f + film[(int64_t)b * film_stride + c];
      const float sh = film[(int64_t)b * film_stride + Ctot + c];
      a *= sc;
      o = o * sc + sh;
    }
    if (pre_add) o += pre_add[(int64_t)b * pre_stride + c] * a;  // norm(h + e) = h*a + (e*a + o)
    A[c] = a;
    Bv[c] = o;
  }
}

// z-slab sharding: this rank's fp64 sums [B][32][2] from its partials (fixed order)
__global__ void __launch_bounds__(1024) gn_reduce_local_kernel(const double* __restrict__ partials, int n_chunks,
                                                               double* __restrict__ sums) {
  const int b = blockIdx.x;
  const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double* ps = partials + (((int64_t)b * 32 + g) * 2) * n_chunks;
  const double* pq = ps + n_chunks;
  double s = 0.0, q = 0.0;
  for (int c = lane; c < n_chunks; c += 32) {
    s += ps[c];
    q += pq[c];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane == 0) {
    sums[((int64_t)b * 32 + g) * 2] = s;
    sums[((int64_t)b * 32 + g) * 2 + 1] = q;
  }
}

// finalize from the all-gathered per-rank sums: gathered[world][B][32][2], summed in rank order
__global__ void __launch_bounds__(1024) gn_finalize_multi_kernel(const double* gathered, int world, int B, int Ctot,
                                                                 int64_t rank_stride, const uint32_t* __restrict__ flags,
                                                                 const uint32_t* __restrict__ seq_ptr, int64_t parity_stride,
                                                                 double inv_count, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, const float* __restrict__ film,
                                                                 int64_t film_stride, const float* __restrict__ pre_add,
                                                                 int64_t pre_stride, float* __restrict__ ab) {
  __shared__ float s_mean[32], s_rstd[32];
  const int b = blockIdx.x;
  pdl_launch_dependents();
  if (flags) {  // peer path: every rank stores its sums into this rank's mailbox and then raises flags[rank] to seq
    const uint32_t seq = *seq_ptr;  // incremented by this rank's push kernel (earlier in the stream)
    gathered += (int64_t)(seq & 1u) * parity_stride;
    if (threadIdx.x < world) {
      const long long t0 = clock64();
      uint32_t v;
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + threadIdx.x) : "memory");
        if (clock64() - t0 > 8000000000LL) {
          printf("ddpm3d slab: statistics of rank %d did not arrive (sequence %u, have %u)\n", (int)threadIdx.x, seq, v);
          __trap();
        }
      } while ((int32_t)(v - seq) < 0);
    }
    __syncthreads();
  }
  if (threadIdx.x < 32) {
    const int g = threadIdx.x;
    double s = 0.0, q = 0.0;
    for (int r = 0; r < world; ++r) {
      const double* p = gathered + (int64_t)r * rank_stride + ((int64_t)b * 32 + g) * 2;
      s += p[0];
      q += p[1];
    }
    const double mean = s * inv_count;
    double var = q * inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[g] = (float)mean;
    s_rstd[g] = (float)(1.0 / sqrt(var + 1e-5));
  }
  __syncthreads();
  const int gpc = Ctot / 32;
  float* A = ab + (int64_t)b * 2 * Ctot;
  float* Bv = A + Ctot;
  for (int c = threadIdx.x; c < Ctot; c += blockDim.x) {
    const int gg = c / gpc;
    float a = s_rstd[gg] * gamma[c];
    float o = beta[c] - s_mean[gg] * a;
    if (film) {
      const float sc = 1.0f + film[(int64_t)b * film_stride + c];
      const float sh = film[(int64_t)b * film_stride + Ctot + c];
      a *= sc;
      o = o * sc + sh;
    }
    if (pre_add) o += pre_add[(int64_t)b * pre_stride + c] * a;
    A[c] = a;
    Bv[c] = o;
  }
}

// Group sums [sum x, sum x^2] of group g of batch element b from the per-CTA channel sums written by the convolution
// epilogues: cs_s[B][P][C_s][2] hold sums of (x - bias_s[c]) and its square (P = one slot per CTA); the bias is folded
// back in fp64: sum x = D1 + n b, sum x^2 = D2 + 2 b D1 + n b^2 with n = voxels per batch element.  128 threads, fixed
// reduction order -> deterministic.  Result in sh[0][0], sh[1][0] (valid for every thread after the call).
__device__ __forceinline__ void chsum_group_sums(const float* __restrict__ cs0, const float* __restrict__ cs1,
                                                 const float* __restrict__ bias0, const float* __restrict__ bias1, int C0,
                                                 int C1, int P0, int P1, double n_vox, int g, int b, double (*sh)[128]) {
  const int tid = threadIdx.x;
  const int Ctot = C0 + C1, gpc = Ctot / 32;
  const int n = (P0 > P1 ? P0 : P1) * gpc;  // (slot, channel-in-group) pairs of this group
  double s = 0.0, q = 0.0;
  // four pairs per thread and pass, all loads issued before the first add: the pass is one L2 round trip, not four
  // (this kernel sits between every convolution and the GroupNorm apply that follows it)
  constexpr int U = 4;
  for (int i0 = tid; i0 < n; i0 += 128 * U) {
    float2 v[U];
    float bc[U];
    int sl[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * 128;
      const int slot = i / gpc, c = g * gpc + i % gpc;
      const bool first = c < C0;
      const int P = first ? P0 : P1;
      // the sources may have different slot counts (the stem: one per CTA)
      const bool ok = i < n && slot < P;
      const float* src = first ? cs0 + (((int64_t)b * P0 + slot) * C0 + c) * 2 : cs1 + (((int64_t)b * P1 + slot) * C1 + (c - C0)) * 2;
      const float* bias = first ? bias0 : bias1;
      v[u] = ok ? *reinterpret_cast<const float2*>(src) : make_float2(0.f, 0.f);
      bc[u] = (ok && bias) ? bias[first ? c : c - C0] : 0.f;
      sl[u] = ok ? slot : -1;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double d1 = (double)v[u].x, d2 = (double)v[u].y, bcd = (double)bc[u];
      s += d1;
      q += d2 + 2.0 * bcd * d1;
      if (sl[u] == 0) { s += n_vox * bcd; q += n_vox * bcd * bcd; }
    }
  }
  // fixed-order reduction: butterfly inside each of the four warps, then the four partials in warp order (one barrier
  // instead of a seven-level shared-memory tree)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if ((tid & 31) == 0) { sh[0][tid >> 5] = s; sh[1][tid >> 5] = q; }
  __syncthreads();
  if (tid == 0) {
    sh[0][0] = ((sh[0][0] + sh[0][1]) + sh[0][2]) + sh[0][3];
    sh[1][0] = ((sh[1][0] + sh[1][1]) + sh[1][2]) + sh[1][3];
  }
  __syncthreads();
}

// finalize from the conv epilogues' channel sums.  grid (32 groups, B), 128 threads.
__global__ void __launch_bounds__(128) gn_finalize_chsum_kernel(const float* __restrict__ cs0, const float* __restrict__ cs1,
                                                                const float* __restrict__ bias0, const float* __restrict__ bias1,
                                                                int C0, int C1, int P0, int P1, double n_vox, double inv_count,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                const float* __restrict__ film, int64_t film_stride,
                                                                float* __restrict__ ab) {
  __shared__ double sh[2][128];
  const int g = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int Ctot = C0 + C1, gpc = Ctot / 32;
  pdl_wait();               // (pdl = 2) this grid was scheduled while the last epilogues of the producing convolution ran
  pdl_launch_dependents();  // the apply kernel may start its prologue now; it waits for this grid before reading `ab`
  // this thread's channel (a group has at most 128 of them up to 4096 channels): fetch its affine parameters now, so
  // that their latency overlaps the reduction instead of following it
  const int c0 = g * gpc + tid;
  const bool mine = tid < gpc;
  float gam = 0.f, bet = 0.f, fsc = 1.f, fsh = 0.f;
  if (mine) {
    gam = gamma[c0];
    bet = beta[c0];
    if (film) {  // h = norm(h) * (1 + scale) + shift
      fsc = 1.0f + film[(int64_t)b * film_stride + c0];
      fsh = film[(int64_t)b * film_stride + Ctot + c0];
    }
  }
  chsum_group_sums(cs0, cs1, bias0, bias1, C0, C1, P0, P1, n_vox, g, b, sh);
  const double mean = sh[0][0] * inv_count;
  double var = sh[1][0] * inv_count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float fmean = (float)mean, rstd = (float)(1.0 / sqrt(var + 1e-5));
  float* A = ab + (int64_t)b * 2 * Ctot;
  float* Bv = A + Ctot;
  if (mine) {
    float a = rstd * gam;
    float o = bet - fmean * a;
    if (film) {
      a *= fsc;
      o = o * fsc + fsh;
    }
    A[c0] = a;
    Bv[c0] = o;
  }
  for (int c = c0 + 128; c < (g + 1) * gpc; c += 128) {  // (more than 4096 channels)
    float a = rstd * gamma[c];
    float o = beta[c] - fmean * a;
    if (film) {
      const float sc = 1.0f + film[(int64_t)b * film_stride + c];
      const float sh2 = film[(int64_t)b * film_stride + Ctot + c];
      a *= sc;
      o = o * sc + sh2;
    }
    A[c] = a;
    Bv[c] = o;
  }
}

template <typename T, typename TO, int N>
__device__ __forceinline__ void gn_put(TO* dst, const float* y) {
  if constexpr (sizeof(TO) == sizeof(T)) {
    Vec<TO> o;
    o.pack(y);
    o.store(dst);
  } else {  // T = bf16, TO = float: two 16-byte stores
#pragma unroll
    for (int k = 0; k < N; k += 4) *reinterpret_cast<float4*>(dst + k) = make_float4(y[k], y[k + 1], y[k + 2], y[k + 3]);
  }
}

// grid (blocks, B); thread -> (row lane, channel vector); the per-(b, c) affine lives in registers.
// Iteration space: output rows for NONE / POOL, input rows for UP.
// The loads of row block i+1 are issued before block i is computed and stored (register double buffer): without it
// all warps of an SM moved through their load / compute / store phases together and the pass reached 4.3 TB/s.
template <typename T, typename TO, int MODE, bool SILU>
__global__ void __launch_bounds__(256, 4) gn_apply_kernel(const T* __restrict__ s0, const T* __restrict__ s1, int C0, int C1, int Z,
                                                       int H, int W, int rows_per_block, const float* __restrict__ ab,
                                                       TO* __restrict__ out, int out_zpad, TO* __restrict__ peer_lo,
                                                       TO* __restrict__ peer_hi) {
  constexpr int N = Vec<T>::N;
  constexpr bool FAST = SILU && sizeof(TO) == 2;  // two-MUFU SiLU
  const int Ctot = C0 + C1;
  const int nvec0 = C0 / N, nvec = Ctot / N;
  const int rpi = blockDim.x / nvec;
  const int v = threadIdx.x % nvec, r = threadIdx.x / nvec;
  const int b = blockIdx.y;
  const int Ho = MODE == RS_POOL ? H / 2 : H, Wo = MODE == RS_POOL ? W / 2 : W;
  const int rows_it = Z * Ho * Wo;        // rows iterated per batch element
  const int rows_in = Z * H * W;
  const int rows_out = MODE == RS_UP ? 4 * rows_in : rows_it;
  const T* base;
  int Csrc;
  if (v < nvec0) { base = s0 + (int64_t)b * rows_in * C0 + v * N; Csrc = C0; }
  else { base = s1 + (int64_t)b * rows_in * C1 + (v - nvec0) * N; Csrc = C1; }
  const int plane_out = rows_out / Z;
  const int64_t obstride = (int64_t)(Z + 2 * out_zpad) * plane_out * Ctot;
  TO* obase = out + ((int64_t)b * (Z + 2 * out_zpad) + out_zpad) * plane_out * Ctot + v * N;
  // z-slab sharding over peer-mapped memory: the first / last plane of this slab is also stored straight into the
  // trailing halo plane of the upper neighbour (peer_lo) / the leading halo plane of the lower one (peer_hi)
  TO* plo = peer_lo ? peer_lo + (int64_t)b * obstride + v * N : nullptr;
  TO* phi = peer_hi ? peer_hi + (int64_t)b * obstride + v * N : nullptr;
  const int last_plane0 = rows_out - plane_out;
  auto put = [&](int orow, const float* y) {
    gn_put<T, TO, N>(obase + (int64_t)orow * Ctot, y);
    if (plo && orow < plane_out) gn_put<T, TO, N>(plo + (int64_t)orow * Ctot, y);
    if (phi && orow >= last_plane0) gn_put<T, TO, N>(phi + (int64_t)(orow - last_plane0) * Ctot, y);
  };
  float A[N], Bv[N];
  pdl_wait();  // launched with programmatic stream serialisation: everything above overlapped the finalize kernel
  pdl_launch_dependents();  // the convolution that follows may set up (barriers, TMEM, descriptors) while this grid drains
  {
    const float* pa = ab + (int64_t)b * 2 * Ctot + v * N;
#pragma unroll
    for (int k = 0; k < N; k += 4) {
      const float4 a4 = *reinterpret_cast<const float4*>(pa + k);
      const float4 b4 = *reinterpret_cast<const float4*>(pa + Ctot + k);
      A[k] = a4.x; A[k + 1] = a4.y; A[k + 2] = a4.z; A[k + 3] = a4.w;
      Bv[k] = b4.x; Bv[k + 1] = b4.y; Bv[k + 2] = b4.z; Bv[k + 3] = b4.w;
    }
  }
  auto act = [&](const float* f, float* y) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const float t = fmaf(f[k], A[k], Bv[k]);
      if (FAST) y[k] = silu_fast(t);
      else y[k] = SILU ? silu_f(t) : t;
    }
  };
  // row blocks are dealt round-robin to the CTAs (see gn_stats_kernel)
  if (MODE == RS_NONE) {
    constexpr int U = 4;
    const int RB = U * rpi;
    const int step = gridDim.x * RB;
    int rb0 = blockIdx.x * RB;
    Vec<T> cur[U], nxt[U];
    if (rb0 + RB <= rows_it) {
#pragma unroll
      for (int u = 0; u < U; ++u) cur[u].load(base + (int64_t)(rb0 + r + u * rpi) * Csrc);
    }
    for (; rb0 < rows_it; rb0 += step) {
      const int row = rb0 + r;
      if (rb0 + RB <= rows_it) {
        const int nb0 = rb0 + step;
        if (nb0 + RB <= rows_it) {  // prefetch the next full block
#pragma unroll
          for (int u = 0; u < U; ++u) nxt[u].load(base + (int64_t)(nb0 + r + u * rpi) * Csrc);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float f[N], y[N];
          cur[u].unpack(f);
          act(f, y);
          put(row + u * rpi, y);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) cur[u] = nxt[u];
      } else {  // ragged last block
        for (int rr = row; rr < rows_it; rr += rpi) {
          Vec<T> a;
          a.load(base + (int64_t)rr * Csrc);
          float f[N], y[N];
          a.unpack(f);
          act(f, y);
          put(rr, y);
        }
      }
    }
  } else if (MODE == RS_POOL) {
    auto load4 = [&](int row, Vec<T>* a) {
      const int wo = row % Wo;
      const int t1 = row / Wo;
      const int ho = t1 % Ho;
      const int z = t1 / Ho;
      const int in_row = (z * H + 2 * ho) * W + 2 * wo;
      a[0].load(base + (int64_t)in_row * Csrc);
      a[1].load(base + (int64_t)(in_row + 1) * Csrc);
      a[2].load(base + (int64_t)(in_row + W) * Csrc);
      a[3].load(base + (int64_t)(in_row + W + 1) * Csrc);
    };
    const int step = gridDim.x * rpi;
    int row = blockIdx.x * rpi + r;
    Vec<T> cur[4], nxt[4];
    if (row < rows_it) load4(row, cur);
    for (; row < rows_it; row += step) {
      if (row + step < rows_it) load4(row + step, nxt);
      float y[N];
#pragma unroll
      for (int k = 0; k < N; ++k) y[k] = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[N], t[N];
        cur[u].unpack(f);
        act(f, t);
#pragma unroll
        for (int k = 0; k < N; ++k) y[k] += t[k];
      }
#pragma unroll
      for (int k = 0; k < N; ++k) y[k] *= 0.25f;
      put(row, y);
#pragma unroll
      for (int u = 0; u < 4; ++u) cur[u] = nxt[u];
    }
  } else {  // RS_UP: one input row -> four output rows
    const int step = gridDim.x * rpi;
    int row = blockIdx.x * rpi + r;
    Vec<T> cur, nxt;
    if (row < rows_it) cur.load(base + (int64_t)row * Csrc);
    for (; row < rows_it; row += step) {
      if (row + step < rows_it) nxt.load(base + (int64_t)(row + step) * Csrc);
      float f[N], y[N];
      cur.unpack(f);
      act(f, y);
      const int w = row % W;
      const int t1 = row / W;
      const int h = t1 % H;
      const int z = t1 / H;
      const int o0 = (z * 2 * H + 2 * h) * (2 * W) + 2 * w;
      put(o0, y);
      put(o0 + 1, y);
      put(o0 + 2 * W, y);
      put(o0 + 2 * W + 1, y);
      cur = nxt;
    }
  }
}

static int gn_threads(int nvec, int* rpi_out) {
  int rpi = 256 / nvec;
  if (rpi < 1) rpi = 1;
  *rpi_out = rpi;
  return nvec * rpi;
}

template <typename T, typename TO>
static int gn_apply_launch(const GnArgs& a, cudaStream_t s) {
  constexpr int N = Vec<T>::N;
  const int Ctot = a.C[0] + a.C[1];
  const int nvec = Ctot / N;
  int rpi;
  const int threads = gn_threads(nvec, &rpi);
  const int Ho = a.resample == RS_POOL ? a.H / 2 : a.H, Wo = a.resample == RS_POOL ? a.W / 2 : a.W;
  const int rows_it = a.Z * Ho * Wo;
  // ~8 CTAs per SM over the whole launch, at least 4 passes of the row lanes per CTA
  int blocks = (int)std::min<int64_t>(ceil_div(rows_it, 4 * rpi), std::max(1, sm_count() * 8 / a.B));
  const int rows_per_block = (int)ceil_div(rows_it, blocks);
  blocks = (int)ceil_div(rows_it, rows_per_block);
  dim3 grid(blocks, a.B);
  const T* s0 = (const T*)a.src[0];
  const T* s1 = (const T*)a.src[1];
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(threads);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // may start while the finalize kernel is still running
  attr[0].val.programmaticStreamSerializationAllowed = a.pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#define GN_LAUNCH(MODE, SILU) \
  DD_CUDA(cudaLaunchKernelEx(&cfg, gn_apply_kernel<T, TO, MODE, SILU>, s0, s1, a.C[0], a.C[1], a.Z, a.H, a.W, rows_per_block, \
                             (const float*)a.ab, (TO*)a.out, a.out_zpad, (TO*)a.peer_halo[0], (TO*)a.peer_halo[1]))
  if (a.silu) {
    if (a.resample == RS_NONE) GN_LAUNCH(RS_NONE, true);
    else if (a.resample == RS_POOL) GN_LAUNCH(RS_POOL, true);
    else GN_LAUNCH(RS_UP, true);
  } else {
    if (a.resample == RS_NONE) GN_LAUNCH(RS_NONE, false);
    else if (a.resample == RS_POOL) GN_LAUNCH(RS_POOL, false);
    else GN_LAUNCH(RS_UP, false);
  }
#undef GN_LAUNCH
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

template <typename T>
static int gn_stats_launch(const GnArgs& a, cudaStream_t s) {
  constexpr int N = Vec<T>::N;
  const int Ctot = a.C[0] + a.C[1];
  const int nvec = Ctot / N;
  int rpi;
  const int threads = gn_threads(nvec, &rpi);
  DD_CHECK(threads <= 256 && threads >= 64, DDPM3D_ERR_ARG, "groupnorm: unsupported channel count");
  const int64_t rows = (int64_t)a.Z * a.H * a.W;
  DD_CHECK(rows * 4 < ((int64_t)1 << 31), DDPM3D_ERR_ARG, "groupnorm: more than 2^29 voxels per batch element");
  const size_t smem = (size_t)rpi * Ctot * 2 * sizeof(double);
  DD_CHECK(smem <= 48 * 1024, DDPM3D_ERR_ARG, "groupnorm: too many channels for the statistics kernel");
  dim3 grid(a.n_chunks, a.B);
  gn_stats_kernel<T><<<grid, threads, smem, s>>>((const T*)a.src[0], (const T*)a.src[1], a.C[0], a.C[1], (int)rows, a.n_chunks,
                                                 a.pre_add, a.pre_stride, a.partials);
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

static int gn_check(const GnArgs& a) {
  const int Ctot = a.C[0] + a.C[1];
  const int N = is_half_dt(a.dt) ? 8 : 4;
  DD_CHECK(Ctot % 32 == 0, DDPM3D_ERR_ARG, "groupnorm: channels must be a multiple of 32");
  DD_CHECK(a.C[0] % N == 0 && a.C[1] % N == 0, DDPM3D_ERR_ARG, "groupnorm: per-source channels must fill 16-byte vectors");
  DD_CHECK(a.resample != RS_POOL || (a.H % 2 == 0 && a.W % 2 == 0), DDPM3D_ERR_ARG, "groupnorm: pool needs even H, W");
  return DDPM3D_OK;
}

static int gn_stats_any(const GnArgs& a, cudaStream_t s) {
  if (a.dt == DDPM3D_BF16) return gn_stats_launch<bf16>(a, s);
  if (a.dt == DDPM3D_FP16) return gn_stats_launch<f16>(a, s);
  return gn_stats_launch<float>(a, s);
}

static int gn_apply_any(const GnArgs& a, cudaStream_t s) {
  if (gn_apply_stream_eligible(a)) return gn_apply_stream(a, s);
  const int dto = a.dt_out < 0 ? a.dt : a.dt_out;  // 16-bit output format (block inputs are fp16, conv operands bf16)
  if (a.dt == DDPM3D_BF16) {
    if (a.out_f32) return gn_apply_launch<bf16, float>(a, s);
    return dto == DDPM3D_FP16 ? gn_apply_launch<bf16, f16>(a, s) : gn_apply_launch<bf16, bf16>(a, s);
  }
  if (a.dt == DDPM3D_FP16) {
    if (a.out_f32) return gn_apply_launch<f16, float>(a, s);
    return dto == DDPM3D_BF16 ? gn_apply_launch<f16, bf16>(a, s) : gn_apply_launch<f16, f16>(a, s);
  }
  return gn_apply_launch<float, float>(a, s);
}

int gn_forward(const GnArgs& a, cudaStream_t s, int* launches) {
  DD_TRY(gn_check(a));
  const int Ctot = a.C[0] + a.C[1];
  DD_TRY(gn_stats_any(a, s));
  const double inv_count = 1.0 / ((double)a.Z * a.H * a.W * (Ctot / 32));
  gn_finalize_kernel<<<a.B, 1024, 0, s>>>(a.partials, a.n_chunks, Ctot, inv_count, a.gamma, a.beta, a.film, a.film_stride,
                                          a.pre_add, a.pre_stride, a.ab);
  DD_CUDA(cudaGetLastError());
  DD_TRY(gn_apply_any(a, s));
  if (launches) *launches += 3;
  return DDPM3D_OK;
}

static int finalize_chsum_launch(const GnArgs& a, double inv_count, cudaStream_t s) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(32, a.B);
  cfg.blockDim = dim3(128);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = a.pdl >= 2 ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DD_CUDA(cudaLaunchKernelEx(&cfg, gn_finalize_chsum_kernel, a.chsum[0], a.chsum[1], a.chsum_bias[0], a.chsum_bias[1], a.C[0], a.C[1],
                             a.chsum_P[0] ? a.chsum_P[0] : chsum_slots(), a.chsum_P[1] ? a.chsum_P[1] : chsum_slots(),
                             (double)a.Z * a.H * a.W, inv_count, a.gamma, a.beta, a.film, a.film_stride, a.ab));
  return DDPM3D_OK;
}

// statistics + per-(b, c) affine only (a.ab); the consumer applies it itself (the fused head, head_tc.cu)
int gn_finalize_only(const GnArgs& a, cudaStream_t s) {
  DD_TRY(gn_check(a));
  const int Ctot = a.C[0] + a.C[1];
  const double inv_count = 1.0 / ((double)a.Z * a.H * a.W * (Ctot / 32));
  if (!a.pre_add && a.chsum[0] && (a.C[1] == 0 || a.chsum[1])) {
    DD_TRY(finalize_chsum_launch(a, inv_count, s));
  } else {
    DD_TRY(gn_stats_any(a, s));
    gn_finalize_kernel<<<a.B, 1024, 0, s>>>(a.partials, a.n_chunks, Ctot, inv_count, a.gamma, a.beta, a.film, a.film_stride,
                                            a.pre_add, a.pre_stride, a.ab);
  }
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

int gn_forward_chsum(const GnArgs& a, cudaStream_t s) {
  DD_TRY(gn_check(a));
  const int Ctot = a.C[0] + a.C[1];
  DD_CHECK(a.chsum[0] && (a.C[1] == 0 || a.chsum[1]) && !a.pre_add, DDPM3D_ERR_STATE, "groupnorm: channel sums missing");
  const double inv_count = 1.0 / ((double)a.Z * a.H * a.W * (Ctot / 32));
  DD_TRY(finalize_chsum_launch(a, inv_count, s));
  return gn_apply_any(a, s);
}

// z-slab sharding with fused statistics: this rank's fp64 group sums [B][32][2] from the conv epilogues' channel sums
__global__ void __launch_bounds__(128) gn_chsum_local_kernel(const float* __restrict__ cs0, const float* __restrict__ cs1,
                                                             const float* __restrict__ bias0, const float* __restrict__ bias1,
                                                             int C0, int C1, int P0, int P1, double n_vox, double* __restrict__ sums) {
  __shared__ double sh[2][128];
  const int g = blockIdx.x, b = blockIdx.y;
  chsum_group_sums(cs0, cs1, bias0, bias1, C0, C1, P0, P1, n_vox, g, b, sh);
  if (threadIdx.x == 0) {
    sums[((int64_t)b * 32 + g) * 2] = sh[0][0];
    sums[((int64_t)b * 32 + g) * 2 + 1] = sh[1][0];
  }
}

int gn_chsum_local(const GnArgs& a, double* sums, cudaStream_t s) {
  DD_TRY(gn_check(a));
  DD_CHECK(a.chsum[0] && (a.C[1] == 0 || a.chsum[1]), DDPM3D_ERR_STATE, "groupnorm: channel sums missing");
  gn_chsum_local_kernel<<<dim3(32, a.B), 128, 0, s>>>(a.chsum[0], a.chsum[1], a.chsum_bias[0], a.chsum_bias[1], a.C[0], a.C[1],
                                                      a.chsum_P[0] ? a.chsum_P[0] : chsum_slots(), a.chsum_P[1] ? a.chsum_P[1] : chsum_slots(), (double)a.Z * a.H * a.W, sums);
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

int gn_stats_local(const GnArgs& a, double* sums, cudaStream_t s) {
  DD_TRY(gn_check(a));
  DD_TRY(gn_stats_any(a, s));
  gn_reduce_local_kernel<<<a.B, 1024, 0, s>>>(a.partials, a.n_chunks, sums);
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

int gn_finalize_apply(const GnArgs& a, cudaStream_t s) {
  const int Ctot = a.C[0] + a.C[1];
  DD_CHECK(a.gathered != nullptr && a.world >= 1, DDPM3D_ERR_STATE, "groupnorm: gathered statistics missing");
  gn_finalize_multi_kernel<<<a.B, 1024, 0, s>>>(a.gathered, a.world, a.B, Ctot, a.gather_stride ? a.gather_stride : (int64_t)a.B * 64,
                                                a.gather_flags, a.gather_seq, a.gather_parity_stride, a.inv_count_global, a.gamma, a.beta, a.film,
                                                a.film_stride, a.pre_add, a.pre_stride, a.ab);
  DD_CUDA(cudaGetLastError());
  return gn_apply_any(a, s);
}

// =================================================================================================
// plain nearest x2 / avg-pool on (H, W)  (Upsample/Downsample without conv, unet.py:81-140)
// =================================================================================================
template <typename T, int MODE>
__global__ void resample_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int Z, int H, int W, int C) {
  constexpr int N = Vec<T>::N;
  const int nvec = C / N;
  const int Ho = MODE == RS_POOL ? H / 2 : H, Wo = MODE == RS_POOL ? W / 2 : W;
  const int64_t total = (int64_t)B * Z * Ho * Wo * nvec;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    const int64_t row = i / nvec;
    if (MODE == RS_POOL) {
      const int wo = (int)(row % Wo);
      const int64_t t1 = row / Wo;
      const int ho = (int)(t1 % Ho);
      const int64_t bz = t1 / Ho;
      const int64_t r = (bz * H + 2 * ho) * W + 2 * wo;
      float acc[N];
#pragma unroll
      for (int k = 0; k < N; ++k) acc[k] = 0.f;
      const int64_t rr[4] = {r, r + 1, r + W, r + W + 1};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        Vec<T> a;
        a.load(in + rr[u] * C + v * N);
        float f[N];
        a.unpack(f);
#pragma unroll
        for (int k = 0; k < N; ++k) acc[k] += f[k];
      }
#pragma unroll
      for (int k = 0; k < N; ++k) acc[k] *= 0.25f;
      Vec<T> o;
      o.pack(acc);
      o.store(out + row * C + v * N);
    } else {
      Vec<T> a;
      a.load(in + row * C + v * N);
      const int w = (int)(row % W);
      const int64_t t1 = row / W;
      const int h = (int)(t1 % H);
      const int64_t bz = t1 / H;
      const int64_t o0 = (bz * (2 * H) + 2 * h) * (2 * W) + 2 * w;
      a.store(out + o0 * C + v * N);
      a.store(out + (o0 + 1) * C + v * N);
      a.store(out + (o0 + 2 * W) * C + v * N);
      a.store(out + (o0 + 2 * W + 1) * C + v * N);
    }
  }
}

int resample_hw(int dt, const void* in, void* out, int B, int Z, int H, int W, int C, int mode, cudaStream_t s) {
  const int N = is_half_dt(dt) ? 8 : 4;
  DD_CHECK(C % N == 0, DDPM3D_ERR_ARG, "resample: channels must fill 16-byte vectors");
  DD_CHECK(mode == RS_POOL || mode == RS_UP, DDPM3D_ERR_ARG, "resample: bad mode");
  const int Ho = mode == RS_POOL ? H / 2 : H, Wo = mode == RS_POOL ? W / 2 : W;
  const int64_t total = (int64_t)B * Z * Ho * Wo * (C / N);
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>(ceil_div(total, threads), sm_count() * 32);
  if (dt == DDPM3D_BF16) {
    if (mode == RS_POOL) resample_kernel<bf16, RS_POOL><<<blocks, threads, 0, s>>>((const bf16*)in, (bf16*)out, B, Z, H, W, C);
    else resample_kernel<bf16, RS_UP><<<blocks, threads, 0, s>>>((const bf16*)in, (bf16*)out, B, Z, H, W, C);
  } else if (dt == DDPM3D_FP16) {
    if (mode == RS_POOL) resample_kernel<f16, RS_POOL><<<blocks, threads, 0, s>>>((const f16*)in, (f16*)out, B, Z, H, W, C);
    else resample_kernel<f16, RS_UP><<<blocks, threads, 0, s>>>((const f16*)in, (f16*)out, B, Z, H, W, C);
  } else {
    if (mode == RS_POOL) resample_kernel<float, RS_POOL><<<blocks, threads, 0, s>>>((const float*)in, (float*)out, B, Z, H, W, C);
    else resample_kernel<float, RS_UP><<<blocks, threads, 0, s>>>((const float*)in, (float*)out, B, Z, H, W, C);
  }
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

// =================================================================================================
// timestep embedding (nn.py:103-121) and the embedding MLPs (unet.py:798-803,199-205)
// =================================================================================================
__device__ __forceinline__ float temb_value(float t, int j, int dim, const float* __restrict__ freqs) {
  const int half = dim / 2;
  if (j >= 2 * half) return 0.f;  // odd dim: zero pad
  const int i = j < half ? j : j - half;
  // th.exp(-math.log(max_period) * arange(half, fp32) / half): the reference evaluates this on the HOST
  // (nn.py:113-115, then .to(device)); when the host table is supplied the angles match it bit for bit
  const float freq = freqs ? freqs[i] : expf(__fdiv_rn(__fmul_rn(-9.210340371976184f, (float)i), (float)half));
  const float ang = __fmul_rn(t, freq);
  return j < half ? cosf(ang) : sinf(ang);
}

__global__ void temb_kernel(const float* __restrict__ t, const float* __restrict__ freqs, float* __restrict__ out, int dim) {
  const int b = blockIdx.x;
  for (int j = threadIdx.x; j < dim; j += blockDim.x) out[(int64_t)b * dim + j] = temb_value(t[b], j, dim, freqs);
}

int timestep_embedding_k(const float* t, const float* freqs, float* out, int B, int dim, cudaStream_t s) {
  temb_kernel<<<B, 128, 0, s>>>(t, freqs, out, dim);
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one CTA per batch element: sinusoid -> Linear -> SiLU -> Linear (+label_emb) -> SiLU
// grid (ceil(ted / TE_ROWS), B): every CTA evaluates the sinusoid and the first Linear in full (128 x 512 weights from L2:
// cheaper than a second launch) and TE_ROWS rows of the second Linear -- one CTA per batch element took 0.11 ms for 1.3 MB of
// weights.  Per-row summation order (lane-strided, then the warp butterfly) is unchanged.
constexpr int TE_ROWS = 16;
__global__ void __launch_bounds__(512) time_embed_kernel(EmbArgs a) {
  extern __shared__ float sm[];  // e0[mc] | h1[ted]
  float* e0 = sm;
  float* h1 = sm + a.model_channels;
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const float t = a.t[b];
  for (int j = threadIdx.x; j < a.model_channels; j += blockDim.x) e0[j] = temb_value(t, j, a.model_channels, a.freqs);
  __syncthreads();
  // first Linear: four rows per warp and pass so that their weight loads are in flight together
  for (int r0 = warp * 4; r0 < a.ted; r0 += nwarp * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane; k < a.model_channels; k += 32) {
      const float e = e0[k];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (r0 + u < a.ted) acc[u] += a.w0[(int64_t)(r0 + u) * a.model_channels + k] * e;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float v = warp_sum(acc[u]);
      if (lane == 0 && r0 + u < a.ted) h1[r0 + u] = silu_f(v + a.b0[r0 + u]);
    }
  }
  __syncthreads();
  const int r = blockIdx.x * TE_ROWS + warp;
  if (warp < TE_ROWS && r < a.ted) {
    const float* w = a.w2 + (int64_t)r * a.ted;
    float acc = 0.f;
    for (int k = lane; k < a.ted; k += 32) acc += w[k] * h1[k];
    acc = warp_sum(acc);
    if (lane == 0) {
      float e = acc + a.b2[r];
      if (a.label_emb) e += a.label_emb[a.y[b] * (int64_t)a.ted + r];  // unet.py:1031-1033
      a.emb_silu[(int64_t)b * a.ted + r] = silu_f(e);
    }
  }
}

// every ResBlock's emb_layers Linear at once: one warp per output row (weights read once for all b)
__global__ void emb_layers_kernel(EmbArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= a.rows_total) return;
  const float* w = a.w_all + (int64_t)warp * a.ted;
  for (int b = 0; b < a.B; ++b) {
    const float* e = a.emb_silu + (int64_t)b * a.ted;
    float acc = 0.f;
    for (int k = lane; k < a.ted; k += 32) acc += w[k] * e[k];
    acc = warp_sum(acc);
    if (lane == 0) a.emb_out[(int64_t)b * a.rows_total + warp] = acc + a.b_all[warp];
  }
}

int embedding_forward(const EmbArgs& a, cudaStream_t s, int* launches) {
  const size_t smem = (size_t)(a.model_channels + a.ted) * sizeof(float);
  time_embed_kernel<<<dim3((unsigned)ceil_div(a.ted, TE_ROWS), a.B), 512, smem, s>>>(a);
  DD_CUDA(cudaGetLastError());
  if (a.rows_total > 0) {
    const int threads = 256;
    const int blocks = (int)ceil_div((int64_t)a.rows_total * 32, threads);
    emb_layers_kernel<<<blocks, threads, 0, s>>>(a);
    DD_CUDA(cudaGetLastError());
  }
  if (launches) *launches += 2;
  return DDPM3D_OK;
}

// =================================================================================================
// p_sample posterior update (gaussian_diffusion.py:262-326, 430-438): one elementwise kernel.
// Every mul/add is an explicit round-to-nearest intrinsic so nothing is contracted into an FMA:
// the result matches the reference's separate torch ops bit for bit except for exp().
// Algorithmic HBM bytes per voxel: read x, eps, v, noise (16 B) + write x_{t-1} (4 B) = 20 B.
// =================================================================================================
struct Philox {
  // Philox4x32-10, counter = (idx_lo, idx_hi, step, 0), key = seed
  __device__ static uint4 rand4(uint64_t seed, uint64_t idx, uint32_t step) {
    uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = step, c3 = 0;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
      const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
      const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
      c0 = n0; c1 = l1; c2 = n2; c3 = l0;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
  __device__ static void normal4(uint64_t seed, uint64_t idx, uint32_t step, float* z) {
    const uint4 r = rand4(seed, idx, step);
    const float u0 = ((float)r.x + 0.5f) * 2.3283064365386963e-10f, u1 = ((float)r.y + 0.5f) * 2.3283064365386963e-10f;
    const float u2 = ((float)r.z + 0.5f) * 2.3283064365386963e-10f, u3 = ((float)r.w + 0.5f) * 2.3283064365386963e-10f;
    const float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
    float s0, c0, s1, c1;
    sincospif(2.0f * u1, &s0, &c0);
    sincospif(2.0f * u3, &s1, &c1);
    z[0] = ra * c0; z[1] = ra * s0; z[2] = rb * c1; z[3] = rb * s1;
  }
};

template <int MEAN, int VAR, bool CLIP>
__global__ void p_sample_update_kernel(UpdateArgs a) {
  const int64_t per_b = (int64_t)a.C * a.n;  // multiple of 4 (checked by the launcher)
  const int64_t total4 = (int64_t)a.B * per_b / 4;
  int exec = 0, cur = -1;
  if (a.step_counter) { cur = a.step_counter[0]; exec = a.step_counter[1]; }
  const float* noise = a.noise ? a.noise + (int64_t)exec * a.noise_step_stride : nullptr;
  constexpr bool LEARNED = (VAR == DDPM3D_VAR_LEARNED || VAR == DDPM3D_VAR_LEARNED_RANGE);
  for (int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < total4; i4 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i4 * 4;
    const int b = (int)(i / per_b);
    const int64_t off = i - (int64_t)b * per_b;  // c*n + j
    int ti = a.t_index ? a.t_index[b] : cur;
    // the reference's table gather raises IndexError for t outside [0, T) (gaussian_diffusion.py:897-910); a kernel
    // cannot raise, so the index is clamped here (like step_from_tensor_kernel) and range-checked by the host mirror
    ti = ti < 0 ? 0 : (ti >= a.T ? a.T - 1 : ti);
    const ddpm3d_step_scalars sc = a.table[ti];
    const int64_t mo_base = (int64_t)b * (LEARNED ? 2 : 1) * per_b + off;
    const float4 x4 = *reinterpret_cast<const float4*>(a.x + i);
    const float4 m4 = *reinterpret_cast<const float4*>(a.model_out + mo_base);
    float4 v4 = make_float4(0, 0, 0, 0);
    if (LEARNED) v4 = *reinterpret_cast<const float4*>(a.model_out + mo_base + per_b);
    float z[4];
    if (noise) {
      const float4 n4 = *reinterpret_cast<const float4*>(noise + i);
      z[0] = n4.x; z[1] = n4.y; z[2] = n4.z; z[3] = n4.w;
    } else {
      const uint64_t gi = a.idx_bstride ? (uint64_t)(((int64_t)b * a.idx_bstride + a.idx_offset + off) >> 2) : (uint64_t)i4;
      Philox::normal4(a.seed, gi, (uint32_t)exec, z);
    }
    const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, ms[4] = {m4.x, m4.y, m4.z, m4.w}, vs[4] = {v4.x, v4.y, v4.z, v4.w};
    float smp[4], x0s[4], mus[4], lvs[4];
    const float mask = ti != 0 ? 1.0f : 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float logvar;
      if (VAR == DDPM3D_VAR_LEARNED) {
        logvar = vs[k];
      } else if (VAR == DDPM3D_VAR_LEARNED_RANGE) {
        const float frac = __fdiv_rn(__fadd_rn(vs[k], 1.0f), 2.0f);
        logvar = __fadd_rn(__fmul_rn(frac, sc.max_log), __fmul_rn(__fsub_rn(1.0f, frac), sc.min_log));
      } else {
        logvar = sc.fixed_log_variance;
      }
      float x0, mu;
      if (MEAN == DDPM3D_MEAN_PREVIOUS_X) {
        x0 = __fsub_rn(__fmul_rn(sc.recip_coef1, ms[k]), __fmul_rn(sc.coef2_over_coef1, xs[k]));
        if (CLIP) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
        mu = ms[k];
      } else {
        if (MEAN == DDPM3D_MEAN_START_X) x0 = ms[k];
        else x0 = __fsub_rn(__fmul_rn(sc.sqrt_recip_alphas_cumprod, xs[k]), __fmul_rn(sc.sqrt_recipm1_alphas_cumprod, ms[k]));
        if (CLIP) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
        mu = __fadd_rn(__fmul_rn(sc.posterior_mean_coef1, x0), __fmul_rn(sc.posterior_mean_coef2, xs[k]));
      }
      if (a.ddim) {
        // eps re-derived from x0 (:345-349), sigma and the Equation-12 mean (:563-580); op order as in the reference
        const float eps = __fdiv_rn(__fsub_rn(__fmul_rn(sc.sqrt_recip_alphas_cumprod, xs[k]), x0), sc.sqrt_recipm1_alphas_cumprod);
        const float ab = sc.alphas_cumprod, abp = sc.alphas_cumprod_prev;
        const float sigma = __fmul_rn(__fmul_rn(a.eta, __fsqrt_rn(__fdiv_rn(__fsub_rn(1.0f, abp), __fsub_rn(1.0f, ab)))),
                                      __fsqrt_rn(__fsub_rn(1.0f, __fdiv_rn(ab, abp))));
        const float mean_pred = __fadd_rn(__fmul_rn(x0, __fsqrt_rn(abp)),
                                          __fmul_rn(__fsqrt_rn(__fsub_rn(__fsub_rn(1.0f, abp), __fmul_rn(sigma, sigma))), eps));
        smp[k] = __fadd_rn(mean_pred, __fmul_rn(__fmul_rn(mask, sigma), z[k]));
      } else {
        const float sd = expf(__fmul_rn(0.5f, logvar));
        smp[k] = __fadd_rn(mu, __fmul_rn(__fmul_rn(mask, sd), z[k]));
      }
      x0s[k] = x0; mus[k] = mu; lvs[k] = logvar;
    }
    *reinterpret_cast<float4*>(a.sample + i) = make_float4(smp[0], smp[1], smp[2], smp[3]);
    if (a.pred_xstart) *reinterpret_cast<float4*>(a.pred_xstart + i) = make_float4(x0s[0], x0s[1], x0s[2], x0s[3]);
    if (a.mean) *reinterpret_cast<float4*>(a.mean + i) = make_float4(mus[0], mus[1], mus[2], mus[3]);
    if (a.log_variance) *reinterpret_cast<float4*>(a.log_variance + i) = make_float4(lvs[0], lvs[1], lvs[2], lvs[3]);
  }
}

template <int MEAN, int VAR>
static void update_launch(const UpdateArgs& a, int blocks, int threads, cudaStream_t s) {
  if (a.clip) p_sample_update_kernel<MEAN, VAR, true><<<blocks, threads, 0, s>>>(a);
  else p_sample_update_kernel<MEAN, VAR, false><<<blocks, threads, 0, s>>>(a);
}
template <int MEAN>
static int update_dispatch_var(const UpdateArgs& a, int blocks, int threads, cudaStream_t s) {
  switch (a.var_type) {
    case DDPM3D_VAR_LEARNED: update_launch<MEAN, DDPM3D_VAR_LEARNED>(a, blocks, threads, s); break;
    case DDPM3D_VAR_LEARNED_RANGE: update_launch<MEAN, DDPM3D_VAR_LEARNED_RANGE>(a, blocks, threads, s); break;
    case DDPM3D_VAR_FIXED_SMALL:
    case DDPM3D_VAR_FIXED_LARGE: update_launch<MEAN, DDPM3D_VAR_FIXED_SMALL>(a, blocks, threads, s); break;
    default: set_error("p_sample_update: bad var_type"); return DDPM3D_ERR_ARG;
  }
  return DDPM3D_OK;
}

int p_sample_update_k(const UpdateArgs& a, cudaStream_t s) {
  DD_CHECK(((int64_t)a.C * a.n) % 4 == 0, DDPM3D_ERR_ARG, "p_sample_update: C*n_spatial must be a multiple of 4");
  DD_CHECK(a.table != nullptr, DDPM3D_ERR_STATE, "p_sample_update: schedule not set");
  const int64_t total4 = (int64_t)a.B * a.C * a.n / 4;
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>(ceil_div(total4, threads), sm_count() * 16);
  switch (a.mean_type) {
    case DDPM3D_MEAN_PREVIOUS_X: DD_TRY(update_dispatch_var<DDPM3D_MEAN_PREVIOUS_X>(a, blocks, threads, s)); break;
    case DDPM3D_MEAN_START_X: DD_TRY(update_dispatch_var<DDPM3D_MEAN_START_X>(a, blocks, threads, s)); break;
    case DDPM3D_MEAN_EPSILON: DD_TRY(update_dispatch_var<DDPM3D_MEAN_EPSILON>(a, blocks, threads, s)); break;
    default: set_error("p_sample_update: bad mean_type"); return DDPM3D_ERR_ARG;
  }
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

// step bookkeeping for the device-resident loop: counter = {current index i, executed steps k}
__global__ void step_advance_kernel(int32_t* counter, float* t_model, const ddpm3d_step_scalars* table, int B) {
  // one warp: lane 0 reads the counter, the others get it by shuffle (no lane reads after lane 0's write)
  int next = 0;
  if (threadIdx.x == 0) {
    next = counter[0] - 1;
    counter[0] = next;
    counter[1] = counter[1] + 1;
  }
  next = __shfl_sync(0xffffffffu, next, 0);
  if (next >= 0)
    for (int b = threadIdx.x; b < B; b += blockDim.x) t_model[b] = table[next].model_t;
}

__global__ void step_set_kernel(int32_t* counter, float* t_model, const ddpm3d_step_scalars* table, int B, int index, int exec) {
  if (threadIdx.x == 0) { counter[0] = index; counter[1] = exec; }
  for (int b = threadIdx.x; b < B; b += blockDim.x) t_model[b] = table[index].model_t;
}

int step_set_k(int32_t* counter, float* t_model, const ddpm3d_step_scalars* table, int B, int index, int exec, cudaStream_t s) {
  step_set_kernel<<<1, 32, 0, s>>>(counter, t_model, table, B, index, exec);
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

// per-sample step indices given as a device tensor (the public p_sample(model, x, t) signature): no host read-back
__global__ void step_from_tensor_kernel(const int64_t* __restrict__ t, int32_t* __restrict__ t_index, float* __restrict__ t_model,
                                        const ddpm3d_step_scalars* __restrict__ table, int B, int T) {
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int64_t i = t[b];
    i = i < 0 ? 0 : (i >= T ? T - 1 : i);
    t_index[b] = (int32_t)i;
    t_model[b] = table[i].model_t;
  }
}

int step_from_tensor_k(const int64_t* t, int32_t* t_index, float* t_model, const ddpm3d_step_scalars* table, int B, int T,
                       cudaStream_t s) {
  step_from_tensor_kernel<<<1, 32, 0, s>>>(t, t_index, t_model, table, B, T);
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

int step_advance_k(int32_t* counter, float* t_model, const ddpm3d_step_scalars* table, int B, cudaStream_t s) {
  step_advance_kernel<<<1, 32, 0, s>>>(counter, t_model, table, B);
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

}  // namespace ddpm3d
