// GroupNorm apply pass for the large tensors as a TMA streaming kernel (HBM bound).
//
//   y[b][row][c] = SiLU(x[b][row][c] * A[b][c] + B[b][c])       x: one tensor or the virtual channel concat of two
//
// The register-staged apply kernel (elementwise.cu) keeps 64 bytes per thread in flight; with the 32 warps per SM its
// registers allow, that is ~64 KB per SM -- just the bandwidth-delay product of HBM3e, and ncu shows it latency bound
// (4.9 TB/s, SFU 60 %, issue slots 49 %).  Here a producer thread streams 16 KB tiles (64 rows x 128 channels) into a
// 6-stage shared-memory ring with TMA, 256 consumer threads rewrite each tile in place and one of them sends it back
// with a TMA store; two CTAs per SM keep up to 192 KB in flight whatever the register budget.  Rows past the end of a
// batch element are zero-filled on load and clipped on store by the TMA unit.
#include "kernels.h"
#include "tc_ptx.cuh"

namespace ddpm3d {

namespace {

constexpr int GS_ROWS = 64, GS_CC = 128, GS_STAGES = 6;
constexpr int GS_CONSUMERS = 256, GS_THREADS = GS_CONSUMERS + 32;
constexpr int GS_TILE_BYTES = GS_ROWS * GS_CC * 2;  // 16 KB
constexpr int GS_STORES_IN_FLIGHT = 2;

struct GsParams {
  int nsrc, chunks[2], coff[2];  // 128-channel chunks of each source; channel offset of each source in the output
  int B, Ctot, nrb;              // nrb = row blocks per batch element
  const float* ab;               // [B][2][Ctot]
};

template <typename T, typename TO, bool SILU>
__global__ void __launch_bounds__(GS_THREADS, 2)
gn_apply_stream_kernel(const __grid_constant__ CUtensorMap mapIn0, const __grid_constant__ CUtensorMap mapIn1,
                       const __grid_constant__ CUtensorMap mapOut, const GsParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * GS_STAGES];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 127u) & ~127u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bar0 = smem_u32(bars);
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (GS_STAGES + s); };
  if (threadIdx.x == 0) {
    prefetch_tmap(&mapIn0);
    prefetch_tmap(&mapOut);
    if (p.nsrc > 1) prefetch_tmap(&mapIn1);
    for (int s = 0; s < GS_STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const int ncombo = p.chunks[0] + (p.nsrc > 1 ? p.chunks[1] : 0);

  if (threadIdx.x >= GS_CONSUMERS) {
    // ===================================== producer ==========================================
    if (threadIdx.x == GS_CONSUMERS) {
      int stage = 0;
      uint32_t ph = 0;
      pdl_wait();  // with the whole chain programmatic, the convolution that writes the input may still be running
      pdl_launch_dependents();
      for (int combo = 0; combo < ncombo; ++combo) {
        const int s = combo < p.chunks[0] ? 0 : 1, chunk = s ? combo - p.chunks[0] : combo;
        const CUtensorMap* map = s ? &mapIn1 : &mapIn0;
        for (int b = 0; b < p.B; ++b)
          for (int rb = blockIdx.x; rb < p.nrb; rb += gridDim.x) {
            mbar_wait(empty(stage), ph ^ 1);
            mbar_expect_tx(full(stage), GS_TILE_BYTES);
            tma_load_3d(base + stage * GS_TILE_BYTES, map, full(stage), chunk * GS_CC, rb * GS_ROWS, b);
            if (++stage == GS_STAGES) { stage = 0; ph ^= 1; }
          }
      }
    }
    return;
  }
  // ======================================= consumers ===========================================
  const int tid = threadIdx.x;
  const int cv = tid & 15, rl = tid >> 4;  // 16-byte channel vector of the chunk, row lane (16 lanes x 4 rows)
  int stage = 0, k = 0;
  uint32_t ph = 0;
  pdl_wait();  // the finalize kernel that writes `ab` may still be running (programmatic stream serialisation)
  for (int combo = 0; combo < ncombo; ++combo) {
    const int s = combo < p.chunks[0] ? 0 : 1, chunk = s ? combo - p.chunks[0] : combo;
    const int cbase = p.coff[s] + chunk * GS_CC;
    for (int b = 0; b < p.B; ++b) {
      float A[8], Bv[8];
      {
        const float* pa = p.ab + (int64_t)b * 2 * p.Ctot + cbase + cv * 8;
#pragma unroll
        for (int i = 0; i < 8; i += 4) {
          const float4 a4 = *reinterpret_cast<const float4*>(pa + i);
          const float4 b4 = *reinterpret_cast<const float4*>(pa + p.Ctot + i);
          A[i] = a4.x; A[i + 1] = a4.y; A[i + 2] = a4.z; A[i + 3] = a4.w;
          Bv[i] = b4.x; Bv[i + 1] = b4.y; Bv[i + 2] = b4.z; Bv[i + 3] = b4.w;
        }
      }
      for (int rb = blockIdx.x; rb < p.nrb; rb += gridDim.x, ++k) {
        mbar_wait(full(stage), ph);
        uint8_t* tile = base_ptr + stage * GS_TILE_BYTES;
        uint4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = *reinterpret_cast<const uint4*>(tile + ((rl + 16 * j) * GS_CC + cv * 8) * 2);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t w[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float x0, x1;
            unpack2<T>(w[i], x0, x1);
            const float t0 = fmaf(x0, A[2 * i], Bv[2 * i]), t1 = fmaf(x1, A[2 * i + 1], Bv[2 * i + 1]);
            o[i] = pack2<TO>(SILU ? silu_fast(t0) : t0, SILU ? silu_fast(t1) : t1);
          }
          *reinterpret_cast<uint4*>(tile + ((rl + 16 * j) * GS_CC + cv * 8) * 2) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the TMA store
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid == 0) {
          tma_store_3d(&mapOut, base + stage * GS_TILE_BYTES, cbase, rb * GS_ROWS, b);
          bulk_commit();
          // the store issued GS_STORES_IN_FLIGHT items ago has finished reading its stage: hand it back to the producer
          bulk_wait_read<GS_STORES_IN_FLIGHT>();
          if (k >= GS_STORES_IN_FLIGHT) mbar_arrive(empty((k - GS_STORES_IN_FLIGHT) % GS_STAGES));
        }
        if (++stage == GS_STAGES) { stage = 0; ph ^= 1; }
      }
    }
  }
  if (tid == 0) {  // drain: the last stores
    bulk_wait_read<0>();
    for (int j = (k >= GS_STORES_IN_FLIGHT ? k - GS_STORES_IN_FLIGHT : 0); j < k; ++j) mbar_arrive(empty(j % GS_STAGES));
  }
}

template <typename T, typename TO>
int launch_stream(const GnArgs& a, cudaStream_t s) {
  const int Ctot = a.C[0] + a.C[1];
  const int64_t rows = (int64_t)a.Z * a.H * a.W;
  GsParams p{};
  p.nsrc = a.C[1] ? 2 : 1;
  p.chunks[0] = a.C[0] / GS_CC; p.chunks[1] = a.C[1] / GS_CC;
  p.coff[0] = 0; p.coff[1] = a.C[0];
  p.B = a.B; p.Ctot = Ctot;
  p.nrb = (int)ceil_div(rows, GS_ROWS);
  p.ab = a.ab;
  const int dto = a.dt_out < 0 ? a.dt : a.dt_out;
  CUtensorMap in0, in1, out;
  DD_TRY(make_rows_map(&in0, tmap_dtype(a.dt), a.src[0], a.B, rows, a.C[0], GS_CC, GS_ROWS));
  in1 = in0;
  if (a.C[1]) DD_TRY(make_rows_map(&in1, tmap_dtype(a.dt), a.src[1], a.B, rows, a.C[1], GS_CC, GS_ROWS));
  DD_TRY(make_rows_map(&out, tmap_dtype(dto), a.out, a.B, rows, Ctot, GS_CC, GS_ROWS));
  const size_t smem = (size_t)GS_STAGES * GS_TILE_BYTES + 128;
  auto kern = a.silu ? gn_apply_stream_kernel<T, TO, true> : gn_apply_stream_kernel<T, TO, false>;
  static uint64_t configured[2] = {0, 0};
  if (first_use_on_device(&configured[a.silu ? 1 : 0]))
    DD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)std::min<int64_t>(p.nrb, 2 * sm_count()));
  cfg.blockDim = dim3(GS_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = a.pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DD_CUDA(cudaLaunchKernelEx(&cfg, kern, in0, in1, out, p));
  return DDPM3D_OK;
}

}  // namespace

bool gn_apply_stream_eligible(const GnArgs& a) {
  if (!a.stream_allowed || !is_half_dt(a.dt) || a.out_f32 || a.resample != RS_NONE || a.out_zpad || a.peer_halo[0] || a.peer_halo[1])
    return false;
  if (a.C[0] % GS_CC != 0 || a.C[1] % GS_CC != 0) return false;
  const int64_t rows = (int64_t)a.Z * a.H * a.W;
  if (rows >= ((int64_t)1 << 31) - GS_ROWS) return false;
  auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!aligned(a.src[0]) || (a.C[1] && !aligned(a.src[1])) || !aligned(a.out)) return false;
  return rows * (a.C[0] + a.C[1]) * 2 * a.B >= ((int64_t)a.stream_min_mb << 20);
}

int gn_apply_stream(const GnArgs& a, cudaStream_t s) {
  const int dto = a.dt_out < 0 ? a.dt : a.dt_out;
  if (a.dt == DDPM3D_BF16) return dto == DDPM3D_FP16 ? launch_stream<bf16, f16>(a, s) : launch_stream<bf16, bf16>(a, s);
  return dto == DDPM3D_BF16 ? launch_stream<f16, bf16>(a, s) : launch_stream<f16, f16>(a, s);
}

}  // namespace ddpm3d
