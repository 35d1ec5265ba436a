// QKVAttention core (unet.py:328-393) fused on the 5th-generation tensor cores, for 64-wide heads.
//
//   S = (Q K^T) / sqrt(ch)      tcgen05.mma  M = 128 queries, N = 128 keys, K = 64 channels   -> TMEM
//   P = softmax_rows(S)         fp32 in registers (a thread owns a query row: no cross-lane reductions)
//   O = P V                     tcgen05.mma  M = 128 queries, N = 64 channels, K = 128 keys   -> TMEM
//
// One CTA per (batch, head, 128-query tile) walks the key tiles twice: pass 1 accumulates the row maximum and the
// normaliser, pass 2 recomputes S, writes the normalised probabilities as a bf16/fp16 SWIZZLE_128B operand into
// shared memory and accumulates O in TMEM -- so O never needs rescaling and the T x T matrix the reference
// materialises (unet.py:349-353) never exists.  Q / K tiles are TMA boxes of the channels-last qkv tensor; V is
// needed K-major (keys contiguous), so a small kernel transposes it once per call.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+ TMEM allocation), 2-5 = softmax / epilogue.
#include "kernels.h"
#include "tc_ptx.cuh"

#include <type_traits>

namespace ddpm3d {

namespace {

constexpr int AT_THREADS = 192;
constexpr int QT = 128, KT = 128, CH = 64;
constexpr int TILE_BYTES = 128 * 128;  // 128 rows x 64 channels x 2 B

struct AttnParams {
  int B, T, H, C, Tp;
  int qb, Tq;                     // query window [qb, qb + Tq) of the T keys (z-slab sharding); out is [B][Tq][C]
  int qoff, koff, voff, hstride;  // channel offsets of q / k / v for head 0 and the per-head stride inside qkv
  float scale2;                   // 1 / sqrt(ch): the reference scales q and k by ch^-1/4 each
  void* out;
};

// V^T per (batch, head): vt[(b*H + h)*64 + c][t] = qkv[b][t][voff + h*hstride + c]
template <typename T>
__global__ void transpose_v_kernel(const T* __restrict__ qkv, T* __restrict__ vt, AttnParams p) {
  __shared__ T tile[64][66];
  const int bh = blockIdx.y, b = bh / p.H, h = bh - b * p.H;
  const int t0 = blockIdx.x * 64;
  const T* src = qkv + ((size_t)b * p.T) * 3 * p.C + p.voff + h * p.hstride;
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int t = i / 64, c = i - t * 64;
    tile[t][c] = (t0 + t < p.T) ? src[(size_t)(t0 + t) * 3 * p.C + c] : from_f32<T>(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int c = i / 64, t = i - c * 64;
    if (t0 + t < p.Tp) vt[((size_t)bh * 64 + c) * p.Tp + t0 + t] = tile[t][c];
  }
}

template <typename T>
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap mapQKV, const __grid_constant__ CUtensorMap mapVt, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[9];
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = base, k_smem = base + TILE_BYTES, v_smem = base + 2 * TILE_BYTES, p_smem = base + 3 * TILE_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t q_full = bar0, k_full = bar0 + 8, k_empty = bar0 + 16, v_full = bar0 + 24, v_empty = bar0 + 32,
                 s_full = bar0 + 40, s_empty = bar0 + 48, p_full = bar0 + 56, o_full = bar0 + 64;

  const int nq = (p.Tq + QT - 1) / QT, nk = (p.T + KT - 1) / KT;
  int blk = blockIdx.x;
  const int qt = blk % nq; blk /= nq;
  const int h = blk % p.H;
  const int b = blk / p.H;
  const int q0 = p.qb + qt * QT;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapQKV);
    prefetch_tmap(&mapVt);
    mbar_init(q_full, 1); mbar_init(k_full, 1); mbar_init(k_empty, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
    mbar_init(s_full, 1); mbar_init(s_empty, 4); mbar_init(p_full, 4); mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t s_tmem = tmem, o_tmem = tmem + 128;
  constexpr bool IS_BF16 = !std::is_same<T, f16>::value;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      mbar_expect_tx(q_full, TILE_BYTES);
      tma_load_2d(q_smem, &mapQKV, q_full, p.qoff + h * p.hstride, b * p.T + q0);
      uint32_t kph = 0, vph = 0;
      for (int pass = 0; pass < 2; ++pass) {
        for (int kt = 0; kt < nk; ++kt) {
          mbar_wait(k_empty, kph ^ 1);
          mbar_expect_tx(k_full, TILE_BYTES);
          tma_load_2d(k_smem, &mapQKV, k_full, p.koff + h * p.hstride, b * p.T + kt * KT);
          kph ^= 1;
          if (pass == 1) {
            mbar_wait(v_empty, vph ^ 1);
            mbar_expect_tx(v_full, TILE_BYTES);
            // two 64-key chunks of V^T: [64 channel rows][64 keys]
            tma_load_2d(v_smem, &mapVt, v_full, kt * KT, (b * p.H + h) * CH);
            tma_load_2d(v_smem + 64 * 128, &mapVt, v_full, kt * KT + 64, (b * p.H + h) * CH);
            vph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc(128, 128, IS_BF16);
      constexpr uint32_t idesc_pv = make_idesc(128, 64, IS_BF16);
      mbar_wait(q_full, 0);
      uint32_t kph = 0, vph = 0, seph = 0, pph = 0;
      const uint64_t qdesc = make_sw128_desc(q_smem), kdesc = make_sw128_desc(k_smem);
      for (int pass = 0; pass < 2; ++pass) {
        for (int kt = 0; kt < nk; ++kt) {
          mbar_wait(k_full, kph); kph ^= 1;
          mbar_wait(s_empty, seph ^ 1); seph ^= 1;  // the softmax warps have read the previous S
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(s_tmem, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_qk, k != 0);
          umma_commit(s_full);
          umma_commit(k_empty);
          if (pass == 1) {
            mbar_wait(p_full, pph); pph ^= 1;
            mbar_wait(v_full, vph); vph ^= 1;
            tc_fence_after();
#pragma unroll
            for (int kc = 0; kc < 2; ++kc) {
              const uint64_t pd = make_sw128_desc(p_smem + kc * (128 * 128));
              const uint64_t vd = make_sw128_desc(v_smem + kc * (64 * 128));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(o_tmem, pd + (uint64_t)(2 * k), vd + (uint64_t)(2 * k), idesc_pv, (kt | kc | k) != 0);
            }
            umma_commit(v_empty);
          }
        }
      }
      umma_commit(o_full);
    }
  } else {
    // ===================================== softmax / epilogue =================================
    const int sub = warp & 3;
    const int row = sub * 32 + lane;  // query row of this thread
    const uint32_t t_lane = (uint32_t)(sub * 32) << 16;
    float m = -INFINITY, l = 0.f;
    uint32_t sph = 0;
    // pass 1: row maximum and normaliser
    for (int kt = 0; kt < nk; ++kt) {
      mbar_wait(s_full, sph); sph ^= 1;
      tc_fence_after();
      float tmax = -INFINITY;
      float sv[4][32];
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        uint32_t r[32];
        tmem_ld32(s_tmem + t_lane + (uint32_t)(c4 * 32), r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float s = (kt * KT + c4 * 32 + j < p.T) ? __uint_as_float(r[j]) * p.scale2 : -INFINITY;
          sv[c4][j] = s;
          tmax = fmaxf(tmax, s);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);
      const float mn = fmaxf(m, tmax);
      float sum = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4)
#pragma unroll
        for (int j = 0; j < 32; ++j) sum += __expf(sv[c4][j] - mn);
      l = l * __expf(m - mn) + sum;
      m = mn;
    }
    const float inv_l = 1.0f / l;
    // pass 2: normalised probabilities -> smem operand
    for (int kt = 0; kt < nk; ++kt) {
      mbar_wait(s_full, sph); sph ^= 1;
      tc_fence_after();
#pragma unroll 1
      for (int c4 = 0; c4 < 4; ++c4) {
        uint32_t r[32];
        tmem_ld32(s_tmem + t_lane + (uint32_t)(c4 * 32), r);
        tmem_ld_wait();
        float pr[32];
#pragma unroll
        for (int j = 0; j < 32; ++j)
          pr[j] = (kt * KT + c4 * 32 + j < p.T) ? __expf(__uint_as_float(r[j]) * p.scale2 - m) * inv_l : 0.f;
        // K-major SWIZZLE_128B layout by hand: chunk kc = key / 64; 16-byte unit u = (key % 64) / 8 lands at u ^ (row & 7)
        const int kc = c4 >> 1;
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) {
          const int u = (c4 & 1) * 4 + uu;
          const uint32_t addr = p_smem + (uint32_t)kc * (128 * 128) + (uint32_t)row * 128u + (uint32_t)((u ^ (row & 7)) << 4);
          const uint32_t w0 = pack2<T>(pr[8 * uu], pr[8 * uu + 1]), w1 = pack2<T>(pr[8 * uu + 2], pr[8 * uu + 3]);
          const uint32_t w2 = pack2<T>(pr[8 * uu + 4], pr[8 * uu + 5]), w3 = pack2<T>(pr[8 * uu + 6], pr[8 * uu + 7]);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
        }
      }
      fence_proxy_async();  // the MMA (async proxy) reads what these generic-proxy stores wrote
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(s_empty); mbar_arrive(p_full); }
    }
    // epilogue: O -> out[b][t][h*64 + c]
    mbar_wait(o_full, 0);
    tc_fence_after();
    const int tq = q0 - p.qb + row;  // query index inside the window
#pragma unroll 1
    for (int c2 = 0; c2 < 2; ++c2) {
      uint32_t r[32];
      tmem_ld32(o_tmem + t_lane + (uint32_t)(c2 * 32), r);
      tmem_ld_wait();
      if (tq < p.Tq) {
        T* op = (T*)p.out + ((size_t)b * p.Tq + tq) * p.C + h * CH + c2 * 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t w4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            w4[q] = pack2<T>(__uint_as_float(r[8 * i + 2 * q]), __uint_as_float(r[8 * i + 2 * q + 1]));
          *reinterpret_cast<uint4*>(op + 8 * i) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

template <typename T>
int launch_attention(const void* qkv, void* out, void* scratch, const AttnParams& p, cudaStream_t s) {
  const CUtensorMapDataType tdt = std::is_same<T, f16>::value ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  dim3 tg((unsigned)ceil_div(p.Tp, 64), (unsigned)(p.B * p.H));
  transpose_v_kernel<T><<<tg, 256, 0, s>>>((const T*)qkv, (T*)scratch, p);
  DD_CUDA(cudaGetLastError());
  CUtensorMap mapQKV, mapVt;
  // make_w_map(map, dtype, ptr, rows, cols, box_rows): 2-D [rows][cols] tensor, box {64 cols, box_rows}
  DD_TRY(make_w_map(&mapQKV, tdt, qkv, p.B * p.T, 3 * p.C, 128));
  DD_TRY(make_w_map(&mapVt, tdt, scratch, p.B * p.H * CH, p.Tp, 64));
  const size_t smem = (size_t)5 * TILE_BYTES + 1024;
  static uint64_t configured = 0;
  if (first_use_on_device(&configured)) {
    DD_CUDA(cudaFuncSetAttribute(attention_tc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int nq = (int)ceil_div(p.Tq, QT);
  attention_tc_kernel<T><<<p.B * p.H * nq, AT_THREADS, smem, s>>>(mapQKV, mapVt, p);
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

}  // namespace

size_t attention_tc_scratch_bytes(int dt, int B, int T, int C, int heads) {
  if (!is_half_dt(dt) || heads <= 0 || C % heads != 0 || C / heads != CH) return 0;
  const size_t Tp = (size_t)ceil_div(T, 64) * 64;
  return (size_t)B * C * Tp * 2;
}

int attention_tc(int dt, const void* qkv, void* out, int B, int T, int C, int heads, int new_order, void* scratch, cudaStream_t s,
                 int q_begin, int q_count) {
  DD_CHECK(attention_tc_scratch_bytes(dt, B, T, C, heads) > 0 && scratch, DDPM3D_ERR_ARG, "attention_tc: not eligible");
  if (q_count < 0) { q_begin = 0; q_count = T; }
  DD_CHECK(q_begin >= 0 && q_count >= 1 && q_begin + q_count <= T, DDPM3D_ERR_ARG, "attention_tc: query window out of range");
  AttnParams p{};
  p.B = B; p.T = T; p.H = heads; p.C = C;
  p.qb = q_begin; p.Tq = q_count;
  p.Tp = (int)ceil_div(T, 64) * 64;
  if (new_order) { p.qoff = 0; p.koff = C; p.voff = 2 * C; p.hstride = CH; }
  else { p.qoff = 0; p.koff = CH; p.voff = 2 * CH; p.hstride = 3 * CH; }
  p.scale2 = 1.0f / sqrtf((float)CH);
  p.out = out;
  if (dt == DDPM3D_BF16) return launch_attention<bf16>(qkv, out, scratch, p, s);
  return launch_attention<f16>(qkv, out, scratch, p, s);
}

}  // namespace ddpm3d
