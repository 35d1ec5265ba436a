// The steps either side of the sampling loop and the ensemble reduction (C5):
//  * extract_patch:    scripts/test.py:205-230  volume (D,H,W) -> zero-padded (P,P,P) patch, (Z,H,W) order
//  * hann_accumulate:  scripts/test.py:91-137   arr += patch * w ; weight += w   on the (H,W,Z) result
//  * hann_finalize:    scripts/test.py:139      arr / weight where weight > 0
//  * welford_update / welford_merge: voxel-wise running mean / M2 over ensemble samples and the
//    pairwise merge of per-rank partials (Chan et al.), fixed order -> deterministic.
// All HBM-bound elementwise kernels; arithmetic follows numpy's promotion rules (fp32 * fp64 -> fp64,
// stored back to fp32) so the blend is bit-identical to the reference's CPU code.
#include "kernels.h"

namespace ddpm3d {

namespace {

__global__ void extract_patch_kernel(const float* __restrict__ vol, int D, int H, int W, int z0, int h0, int w0, int P,
                                     float* __restrict__ out) {
  const int64_t n = (int64_t)P * P * P;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % P), h = (int)((i / P) % P), z = (int)(i / ((int64_t)P * P));
    const int gz = z0 + z, gh = h0 + h, gw = w0 + w;
    out[i] = (gz < D && gh < H && gw < W) ? vol[((int64_t)gz * H + gh) * W + gw] : 0.f;
  }
}

// patch: (P,P,P) in (Z,H,W) order (the sampler's layout); arr / wsum: (H,W,Z) order (the reference's output layout)
__global__ void hann_accumulate_kernel(const float* __restrict__ patch, const double* __restrict__ win, double win_max, int P,
                                       int D, int H, int W, int z0, int h0, int w0, float* __restrict__ arr,
                                       float* __restrict__ wsum) {
  const int hx = min(P, H - h0), wy = min(P, W - w0), dz = min(P, D - z0);
  const int64_t n = (int64_t)hx * wy * dz;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int z = (int)(i % dz), w = (int)((i / dz) % wy), h = (int)(i / ((int64_t)dz * wy));
    // create_3d_hann_window: outer(outer(h1, h1).flatten(), h1) / max, all fp64 (scripts/test.py:248-262)
    const double wt = __ddiv_rn(__dmul_rn(__dmul_rn(win[h], win[w]), win[z]), win_max);
    const float pv = patch[((int64_t)z * P + h) * P + w];
    const int64_t o = ((int64_t)(h0 + h) * W + (w0 + w)) * D + (z0 + z);
    arr[o] = (float)__dadd_rn((double)arr[o], __dmul_rn((double)pv, wt));
    wsum[o] = (float)__dadd_rn((double)wsum[o], wt);
  }
}

__global__ void hann_finalize_kernel(float* __restrict__ arr, const float* __restrict__ wsum, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (wsum[i] > 0.f) arr[i] = __fdiv_rn(arr[i], wsum[i]);
}

__global__ void welford_update_kernel(float* __restrict__ mean, float* __restrict__ m2, const float* __restrict__ x,
                                      float count_after, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float d = v - mean[i];
    const float m = mean[i] + d / count_after;
    mean[i] = m;
    m2[i] = fmaf(d, v - m, m2[i]);
  }
}

__global__ void welford_merge_kernel(float* __restrict__ mean_a, float* __restrict__ m2_a, float na,
                                     const float* __restrict__ mean_b, const float* __restrict__ m2_b, float nb, int64_t n) {
  const float nt = na + nb;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = mean_b[i] - mean_a[i];
    mean_a[i] = mean_a[i] + d * (nb / nt);
    m2_a[i] = m2_a[i] + m2_b[i] + d * d * (na * nb / nt);
  }
}

int grid_for(int64_t n) { return (int)std::min<int64_t>(ceil_div(n, 256), sm_count() * 16); }

}  // namespace
}  // namespace ddpm3d

using namespace ddpm3d;

extern "C" {

int ddpm3d_k_extract_patch(const float* vol, int D, int H, int W, int z0, int h0, int w0, int P, float* out, void* stream) {
  DD_CHECK(vol && out && P > 0 && z0 >= 0 && h0 >= 0 && w0 >= 0, DDPM3D_ERR_ARG, "k_extract_patch: bad argument");
  extract_patch_kernel<<<grid_for((int64_t)P * P * P), 256, 0, (cudaStream_t)stream>>>(vol, D, H, W, z0, h0, w0, P, out);
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

int ddpm3d_k_hann_accumulate(const float* patch, const double* window, double window_max, int P, int D, int H, int W, int z0,
                             int h0, int w0, float* arr, float* wsum, void* stream) {
  DD_CHECK(patch && window && arr && wsum && P > 0, DDPM3D_ERR_ARG, "k_hann_accumulate: bad argument");
  DD_CHECK(z0 >= 0 && h0 >= 0 && w0 >= 0 && z0 < D && h0 < H && w0 < W, DDPM3D_ERR_ARG, "k_hann_accumulate: start outside the volume");
  hann_accumulate_kernel<<<grid_for((int64_t)P * P * P), 256, 0, (cudaStream_t)stream>>>(patch, window, window_max, P, D, H, W, z0,
                                                                                         h0, w0, arr, wsum);
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

int ddpm3d_k_hann_finalize(float* arr, const float* wsum, int64_t n, void* stream) {
  DD_CHECK(arr && wsum && n >= 0, DDPM3D_ERR_ARG, "k_hann_finalize: bad argument");
  hann_finalize_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(arr, wsum, n);
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

int ddpm3d_k_welford_update(float* mean, float* m2, const float* x, int count_after, int64_t n, void* stream) {
  DD_CHECK(mean && m2 && x && count_after >= 1, DDPM3D_ERR_ARG, "k_welford_update: bad argument");
  welford_update_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(mean, m2, x, (float)count_after, n);
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

int ddpm3d_k_welford_merge(float* mean_a, float* m2_a, int na, const float* mean_b, const float* m2_b, int nb, int64_t n,
                           void* stream) {
  DD_CHECK(mean_a && m2_a && mean_b && m2_b && na >= 0 && nb >= 0 && na + nb >= 1, DDPM3D_ERR_ARG, "k_welford_merge: bad argument");
  welford_merge_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(mean_a, m2_a, (float)na, mean_b, m2_b, (float)nb, n);
  DD_CUDA(cudaGetLastError());
  return DDPM3D_OK;
}

}  // extern "C"
