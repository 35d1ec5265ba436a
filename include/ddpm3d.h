/*
 * ddpm3d.h -- C ABI of the B200-native 3D-DDPM sampling hot path.
 *
 * The reference (Zachary-Luk/3D-Denoising-Diffusion-Model) is pure Python/PyTorch and has no
 * FFI layer; its boundary for this path is the Python API used by scripts/test.py:26-35,61-69.
 * Each entry point below names the reference interface it replaces (file:line relative to the
 * reference root).  The Python host mirror in `3d-denoising-diffusion-model_b200/` binds these
 * with ctypes and re-exposes the reference's own names.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary;
 *   - every function returns 0 on success, <0 on error; ddpm3d_last_error() gives the message
 *     (thread-local).  No exception crosses the ABI;
 *   - all device work is enqueued on the caller-supplied cudaStream_t (passed as void*); calls
 *     are asynchronous unless stated;
 *   - the caller owns every input/output buffer; the library owns its packed-weight arena and
 *     its activation workspace;
 *   - volumes are the reference's NCDHW fp32 tensors, (B, C, Z, H, W) with Z the (never strided)
 *     long body axis; `t` is the ORIGINAL-numbering timestep the model sees (after
 *     respace.py:123-128), as fp32 (nn.py:117 casts it to float anyway);
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     DDPM3D_ERR_CUDA.
 */
#ifndef DDPM3D_H_
#define DDPM3D_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDPM3D_ABI_VERSION 5

#define DDPM3D_OK 0
#define DDPM3D_ERR_ARG (-1)     /* bad argument / unsupported configuration */
#define DDPM3D_ERR_CUDA (-2)    /* CUDA runtime / driver error (incl. no device) */
#define DDPM3D_ERR_STATE (-3)   /* call order violated (e.g. forward before finalize) */
#define DDPM3D_ERR_MISSING (-4) /* state_dict key missing / unknown */

/* precision of the UNet torso (unet.py:1003-1013 convert_to_fp16 -> here bf16 on tcgen05) */
#define DDPM3D_FP32 0
#define DDPM3D_BF16 1 /* bf16 tensor-core operands (3x3x3 / qkv / proj weights, GroupNorm outputs); the tensors that are NOT
                         operands of the dense contraction -- ResBlock inputs / outputs and the tensor between a block's two
                         convs, read only by GroupNorm, the 1x1x1 skip path and residual adds -- are stored as fp16 (same
                         bytes, 3 more mantissa bits: eps max-rel x0.65) */
#define DDPM3D_FP16 2 /* the reference's own torso dtype; same tcgen05 rate as bf16, 3 more mantissa bits */
#define DDPM3D_BF16_STRICT 3 /* every 16-bit tensor bf16 (the round-1 behaviour; kept for comparison) */

/* gaussian_diffusion.py:65-72 ModelMeanType */
#define DDPM3D_MEAN_PREVIOUS_X 0
#define DDPM3D_MEAN_START_X 1
#define DDPM3D_MEAN_EPSILON 2
/* gaussian_diffusion.py:75-86 ModelVarType */
#define DDPM3D_VAR_LEARNED 0
#define DDPM3D_VAR_FIXED_SMALL 1
#define DDPM3D_VAR_FIXED_LARGE 2
#define DDPM3D_VAR_LEARNED_RANGE 3

#define DDPM3D_MAX_LEVELS 8

typedef struct ddpm3d_ctx ddpm3d_ctx;

/* Arguments of UNetModel_noatt.__init__ (guided_diffusion/unet.py:751-772) as filled by
 * sr_create_model (script_util.py:334-450). */
typedef struct ddpm3d_config {
  int32_t image_size;      /* large_size; informational */
  int32_t in_channels;     /* channels of x (1 for the PET model); a conditional model doubles it for the low_res
                              concat (unet.py:1683, :1666-1673) unless `unconditional` is set */
  int32_t model_channels;  /* num_channels */
  int32_t out_channels;    /* 2 if learn_sigma else 1 */
  int32_t num_res_blocks;
  int32_t n_levels;
  int32_t channel_mult[DDPM3D_MAX_LEVELS];
  int32_t n_attention_ds;
  int32_t attention_ds[DDPM3D_MAX_LEVELS]; /* large_size // res, script_util.py:363-365 */
  int32_t num_classes;     /* 0 = not class conditional */
  int32_t num_heads;
  int32_t num_head_channels; /* -1 = use num_heads */
  int32_t num_heads_upsample; /* -1 = num_heads */
  int32_t use_scale_shift_norm;
  int32_t resblock_updown;
  int32_t use_new_attention_order;
  int32_t precision;       /* DDPM3D_FP32 | DDPM3D_BF16 | DDPM3D_FP16 | DDPM3D_BF16_STRICT */
  /* ABI 3: the other model classes of unet.py.  All-zero = SuperResModel_noatt with dims=3 (the live model). */
  int32_t dims;            /* 0 or 3: Conv3d, (1,2,2) resampling; 2: the 2-D UNetModel (unet.py:396-716): weights
                              are [Cout][Cin][3][3], activations (B,C,H,W) are passed with Z = 1 */
  int32_t middle_attention; /* 1: UNetModel's AttentionBlock between the two middle ResBlocks (unet.py:548-555) */
  int32_t unconditional;   /* 1: UNetModel.forward(x, t, y): no low_res concat, low_res must be NULL */
} ddpm3d_config;

/* Per-timestep scalars of the respaced process, already rounded to fp32 exactly as
 * _extract_into_tensor's `.float()` does (gaussian_diffusion.py:897-910).  Computed on the host in
 * fp64 by the Python mirror (bit-exact restatement of gaussian_diffusion.py:118-169 and
 * respace.py:72-86). */
typedef struct ddpm3d_step_scalars {
  float model_t;                       /* timestep_map[i] (x 1000/T0 if rescale_timesteps) */
  float sqrt_recip_alphas_cumprod;     /* :328-333 */
  float sqrt_recipm1_alphas_cumprod;
  float posterior_mean_coef1;          /* :208-230 */
  float posterior_mean_coef2;
  float min_log;                       /* posterior_log_variance_clipped[i]   (:270-272) */
  float max_log;                       /* log(betas[i])                         (:273)   */
  float fixed_variance;                /* FIXED_SMALL / FIXED_LARGE tables      (:279-293) */
  float fixed_log_variance;
  float recip_coef1;                   /* 1/posterior_mean_coef1                (:335-343) */
  float coef2_over_coef1;
  float alphas_cumprod;                /* DDIM (:567-568) */
  float alphas_cumprod_prev;
  float pad_[3];
} ddpm3d_step_scalars;

const char* ddpm3d_last_error(void);
int ddpm3d_abi_version(void);

/* ---- model lifetime: replaces sr_create_model (script_util.py:334-450) -------------------- */
int ddpm3d_create(const ddpm3d_config* cfg, ddpm3d_ctx** out);
void ddpm3d_destroy(ddpm3d_ctx* ctx);

/* The state_dict contract: key i, its shape.  Replaces nn.Module.state_dict()
 * key enumeration of UNetModel_noatt (unet.py:797-997). */
int ddpm3d_param_count(const ddpm3d_ctx* ctx);
int ddpm3d_param_info(const ddpm3d_ctx* ctx, int index, const char** key, int64_t shape[8], int* ndim);

/* Replaces model.load_state_dict(...) + model.to(dev) + model.convert_to_fp16()
 * (scripts/test.py:29-35).  `data` is fp32, host or device memory, laid out as the reference
 * tensor (e.g. [Cout,Cin,3,3,3]).  finalize packs everything into the device weight arena
 * (bf16 [Cout][27*Cin (+skip Cin)] K-major for the tcgen05 path) on CUDA device `device`; it
 * synchronises. */
int ddpm3d_load_tensor(ddpm3d_ctx* ctx, const char* key, const float* data, const int64_t* shape, int ndim);
int ddpm3d_finalize_weights(ddpm3d_ctx* ctx, int device);

/* nn.py:113-115 evaluates freqs = exp(-ln(10000) * arange(half) / half) on the HOST in fp32 and moves the
 * table to the device; handing the library that very table (host or device pointer, n = model_channels/2)
 * makes the sinusoid arguments bit-identical to the reference's.  Optional: without it the device computes
 * the frequencies itself (<= 1 ulp apart, i.e. <= 6e-5 absolute on the embedding at t = 999). */
int ddpm3d_set_timestep_freqs(ddpm3d_ctx* ctx, const float* freqs, int n);

/* Bytes of activation workspace the library allocates for a (B,Z,H,W) problem. */
int64_t ddpm3d_workspace_bytes(ddpm3d_ctx* ctx, int B, int Z, int H, int W);

/* ---- one UNet evaluation: replaces SuperResModel_noatt.forward (unet.py:1687-1694) ---------
 * x, low_res: device fp32 (B,in_channels,Z,H,W) (low_res NULL for an unconditional model; Z = 1 when dims = 2);
 * t: device fp32 (B); y: device int64 (B) or NULL; out: device fp32 (B,out_channels,Z,H,W). */
int ddpm3d_unet_forward(ddpm3d_ctx* ctx, const float* x, const float* low_res, const float* t,
                        const int64_t* y, float* out, int B, int Z, int H, int W, void* stream);

/* ---- sampler: replaces GaussianDiffusion/SpacedDiffusion (gaussian_diffusion.py:232-535) ---
 * set_schedule uploads the T-entry scalar table (host pointer) and the mean/var modes. */
int ddpm3d_set_schedule(ddpm3d_ctx* ctx, const ddpm3d_step_scalars* table, int T, int mean_type, int var_type);

/* The elementwise half of p_sample (gaussian_diffusion.py:262-326,430-438): from the raw model
 * output to x_{t-1}.  x, noise, sample, pred_xstart, mean, log_variance: device fp32 (B,C,n_spatial);
 * model_out: (B,2C,n) for learned variance else (B,C,n); t_index: device int32 (B) indices into the
 * schedule table.  pred_xstart / mean / log_variance may be NULL.  One kernel launch. */
int ddpm3d_p_sample_update(ddpm3d_ctx* ctx, const float* x, const float* model_out, const float* noise,
                           const int32_t* t_index, int clip_denoised, float* sample, float* pred_xstart,
                           float* mean, float* log_variance, int B, int C, int64_t n_spatial, void* stream);

/* DDIM (gaussian_diffusion.py:537-585): kind 0 = ancestral DDPM update (default), 1 = DDIM with `eta`.  Applies to
 * ddpm3d_p_sample_update / ddpm3d_p_sample / ddpm3d_sample_loop (= ddim_sample / ddim_sample_loop, :625-707).  The
 * DDIM update has no exp(): with correctly rounded sqrt / div it matches the reference's fp32 torch ops to the bit on CUDA. */
int ddpm3d_set_sampler(ddpm3d_ctx* ctx, int kind, float eta);

/* p_sample (gaussian_diffusion.py:395-439) = UNet + update for step index `i` (same for the whole
 * batch, as p_sample_loop_progressive :522-525 issues it). */
int ddpm3d_p_sample(ddpm3d_ctx* ctx, const float* x, const float* low_res, const int64_t* y, const float* noise,
                    int step_index, int clip_denoised, float* sample, float* pred_xstart,
                    int B, int Z, int H, int W, void* stream);

/* Same with the step indices as the device int64 tensor `t` (B) of the public signature p_sample(model, x, t, ...),
 * possibly different per sample: nothing is read back to the host. */
int ddpm3d_p_sample_t(ddpm3d_ctx* ctx, const float* x, const float* low_res, const int64_t* y, const float* noise,
                      const int64_t* t, int clip_denoised, float* sample, float* pred_xstart, int B, int Z, int H, int W,
                      void* stream);

/* p_sample_loop (gaussian_diffusion.py:441-485): runs steps i = T-1 ... T-n_steps (n_steps <= 0: all T)
 * entirely on the device.  noise: device fp32 [n_steps][B*Z*H*W] consumed in execution order
 * (replaces th.randn_like, :430), or NULL to draw it in-kernel from Philox4x32-10 with `seed`.
 * x_T, low_res, out: device fp32 (B,1,Z,H,W).  out may alias x_T. */
int ddpm3d_sample_loop(ddpm3d_ctx* ctx, const float* x_T, const float* low_res, const int64_t* y,
                       const float* noise, uint64_t seed, int clip_denoised, int n_steps, float* out,
                       int B, int Z, int H, int W, void* stream);

/* ---- one large volume as z-slabs over ranks (not in the reference) ---------------
 * Z is never strided by the network (unet.py:102-105,129), so every 3x3x3 conv needs exactly one halo plane
 * from each z-neighbour and GroupNorm needs the global per-group sums.  After ddpm3d_set_comm (world > 1) the
 * forward / sampler entry points take THIS RANK'S slab (B,1,Zl,H,W): conv-input tensors carry two halo planes
 * filled by NCCL send/recv, GroupNorm all-gathers fp64 partial sums and adds them in rank order (deterministic).
 * comm_unique_id: rank 0 creates the 128-byte NCCL id, the host broadcasts it (torch.distributed), every
 * rank calls set_comm.  set_slab(z_begin, z_total): where this rank's slab sits in the global volume (GroupNorm
 * count, Philox counters); it switches the sharded path ON, and every following forward / sampler call must be a
 * proper slab (Zl < z_total) of that volume.  set_slab(0, 0) switches it OFF again (independent patches, ensemble
 * samples on the same context).  Requires resblock_updown=1.  Attention levels are supported with equal slabs in
 * rank order (z_total = world * Zl): queries stay local, the qkv rows are all-gathered along T. */
int ddpm3d_comm_unique_id(void* out128);
int ddpm3d_set_comm(ddpm3d_ctx* ctx, const void* id128, int rank, int world);
int ddpm3d_set_slab(ddpm3d_ctx* ctx, int z_begin, int z_total);

/* ---- knobs and introspection ------------------------------------------------------------------ */
/* "cuda_graph" (0/1, default 1): replay one captured graph per UNet evaluation;
 * "conv_path" (0 = auto, 1 = force the generic CUDA-core kernel everywhere, 2 = same as 0);
 * "profile" (0/1/2): 1 = record a CUDA-event pair around every operator, launched eagerly (graphs off: the short kernels
 * then include launch gaps); 2 = the step is captured as usual and the brackets become event-record nodes of the graph,
 * so ddpm3d_profile_read returns the operator times of the LAST replay of the captured graph;
 * "fuse_stats" (0/1, default 1): GroupNorm statistics come from per-channel sums accumulated in the producing
 * convolution's epilogue instead of a separate pass over the tensor;
 * "split_k" (0/1, default 1): convolutions with too few tiles to fill the GPU deal their k-steps evenly to the CTAs
 * (stream-K: fp32 partials, deterministic fix-up pass);
 * "strip" (0/1/2, default 2): 3x3x3 layers stage the A operand once per (dz, channel chunk) and address the 9
 * in-plane taps through row-shifted descriptors, with swapped MMA operands (M = channels, N = up to 256 voxels):
 * half the L2 traffic, -9 % on the network step.  1 = only planes >= 24 wide with >= 2 tiles per SM, 2 = also 12..23-wide
 * planes, two z-planes per tile sharing every weight tile, where that beats the stream-K brick kernel (the 384-channel
 * 12 x 12 layers of the shipped network); "strip_w" (3..8, default 4): most stages of its weight ring;
 * "cluster" (0/1, default 0): 2-CTA clusters multicast the weight tile (neutral);
 * "fold_identity" (0/1, default 1): identity skips enter the second conv of a ResBlock as a unit-weight 1x1x1 source;
 * "stem_tc" (0/1, default 1): the Cin == 2 stem runs as one tcgen05 tile per 128 voxels in the 16-bit modes;
 * "head_v2" (0/1, default 1): the fp32 head conv uses 32x16x4 bricks with cp.async double-buffered channel stages;
 * "slab_p2p" (0/1, default 1; set before ddpm3d_set_comm): z-slab sharding exchanges halo planes and GroupNorm sums through
 * peer-mapped memory (CUDA IPC over NVLink: the producing kernels store into the neighbours' halo planes, sequence-numbered
 * flags order the accesses) instead of NCCL send/recv and all-gather; needs equal slabs and peer access, else NCCL is used;
 * "pdl" (0/1/2, default 2): programmatic dependent launch (griddepcontrol.wait / launch_dependents).  1 = the GroupNorm
 * apply kernel may be scheduled while the finalize kernel before it runs; 2 = the whole convolution / GroupNorm chain:
 * the finalize kernel is scheduled while the producing convolution runs, and a tcgen05 convolution sets up its barriers,
 * TMEM and descriptors while the apply pass before it drains (every thread waits for the grid before it before touching
 * global memory);
 * "head_tc" (0/1, default 1): 16-bit modes with 64 / 128 model channels: out.0 GroupNorm apply + SiLU + out.2 conv as one
 * tcgen05 kernel (the contraction over channels once per voxel, the 27 taps as a shifted sum); 0 = GroupNorm pass writing
 * fp32 + the CUDA-core head. */
int ddpm3d_set_option(ddpm3d_ctx* ctx, const char* name, int64_t value);
/* Kernel launches enqueued by this ctx since creation. */
int64_t ddpm3d_launch_count(const ddpm3d_ctx* ctx);
/* After a profiled call: synchronises and writes up to `cap` records; returns the record count.
 * `kind`: 0 conv-tcgen05, 1 conv-simt, 2 gn-stats, 3 gn-finalize, 4 gn-apply, 5 embedding, 6 update,
 * 7 attention, 8 pack/resample/misc, 9 conv-small (stem / head direct convolutions), 10 halo exchange, 11 GroupNorm
 * statistics exchange, 12 an empty bracket (what two back-to-back event records cost in this launch mode).  `work` = algorithmic flops (conv, attention) or bytes (others). */
typedef struct ddpm3d_prof_record { int32_t kind; int32_t pad_; float ms; float pad2_; double work; } ddpm3d_prof_record;
int ddpm3d_profile_read(ddpm3d_ctx* ctx, ddpm3d_prof_record* out, int cap);

/* ---- single kernels on channels-last device buffers, exported for unit tests ------------------
 * dtype: DDPM3D_FP32 (float), DDPM3D_BF16 (__nv_bfloat16) or DDPM3D_FP16 (__half) for activations and conv weights.
 * Activations are [B][Z][H][W][C] (NDHWC).  These replace the torch ops behind nn.py:17-32. */

/* 3x3x3 (taps=27), in-plane 3x3 (taps=9: the Conv2d of a dims = 2 network on one-plane volumes) or 1x1x1 (taps=1) "same"
 * convolution, stride (1,s,s) (nn.py:22-32; call sites unet.py:185,211,219,222).  w: [Cout][taps*Cin] with
 * k = tap*Cin + ci, tap = (dz*3+dh)*3+dw (taps=9: tap = dh*3+dw);
 * bias fp32 [Cout]; residual (optional, [B][Z][Ho][Wo][Cout]) is added in the epilogue.
 * path: 1 SIMT, 2 tcgen05 (bf16 / fp16, Cin%64==0, Cout%64==0; s==2 through element-strided TMA boxes), 3 / 4 the Cin==2 stem kernels
 * (3: tcgen05 tile per 128 voxels in the 16-bit modes, 4: CUDA cores), 5 / 6 tcgen05 with the strip variant off / restricted
 * to the large layers ("strip" option levels 0 / 1), 7 the strip variant with a 4-stage weight ring. */
int ddpm3d_k_conv3d(int dtype, int path, const void* in, const void* w, const float* bias, const void* residual,
                    void* out, int B, int Z, int H, int W, int Cin, int Cout, int taps, int stride_hw, void* stream);

/* Which tcgen05 kernel and tiling ddpm3d_k_conv3d (path 2) / the engine would pick for a 3x3x3 (taps 27), 3x3 (9) or
 * 1x1x1 (1) layer on a device with `sms` SMs (0 = the current device): the planning logic of csrc/conv_tc.cu as plain host
 * arithmetic -- nothing is launched and with sms > 0 no device is needed (the CPU test suite pins the plans of the
 * shipped network).  extra_channels: channels of a folded 1x1x1 skip source (0 = none); split_k / strip: the options of
 * the same names.  out8 = {kind: 0 not eligible, 1 brick, 2 brick + stream-K, 3 strip; MT | z-planes per strip tile;
 * BN | positions per strip tile (MMA N); tiles; grid; split-K slots | weight-ring stages; k-steps | macro steps; strip
 * box rows}. */
int ddpm3d_k_conv_plan(int dtype, int B, int Z, int H, int W, int Cin, int Cout, int taps, int extra_channels, int split_k,
                       int strip, int sms, int32_t* out8);

/* Hardware probe used by the tests / design work: D = A[shift : shift+128, :] * I on tcgen05 with the A operand
 * descriptor starting `shift` rows into a [rows][64] bf16 SWIZZLE_128B tile (mode 0: base-offset field 0, mode 1:
 * base-offset = (start >> 7) & 7).  a: device bf16 [rows][64], ident: device bf16 64x64 identity, out: fp32 [128][64]. */
int ddpm3d_k_probe_rowshift(const void* a, int rows, const void* ident, int shift, int mode, float* out, void* stream);

/* GroupNorm32(32,C) (+ optional per-(b,c) FiLM scale/shift, unet.py:248-252) (+ optional SiLU)
 * (+ optional AvgPool (1,2,2) / nearest x2 on (H,W) of the result, unet.py:81-140):
 * resample 0 none, 1 pool, 2 upsample.  gamma/beta fp32 [C]; film fp32 [B][2C] (scale then shift) or NULL. */
int ddpm3d_k_groupnorm(int dtype, const void* in, const float* gamma, const float* beta, const float* film,
                       int silu, int resample, void* out, int B, int Z, int H, int W, int C, void* stream);

/* The fused statistics path of a ResBlock (unet.py:245-247: conv -> GroupNorm32): the tcgen05 convolution accumulates
 * per-channel sums of its own output in the epilogue and GroupNorm (no FiLM, no SiLU) normalises from them without a
 * statistics pass.  conv_out / gn_out: [B][Z][H][W][Cout] of `dtype` (16-bit).  Unit-test entry point. */
int ddpm3d_k_conv3d_gn(int dtype, const void* in, const void* w, const float* bias, const float* gamma, const float* beta,
                       void* conv_out, void* gn_out, int B, int Z, int H, int W, int Cin, int Cout, void* stream);

/* timestep_embedding (nn.py:103-121): t fp32 [B] -> out fp32 [B][dim] (cos block first).  freqs: device fp32
 * [dim/2] host-computed table (see ddpm3d_set_timestep_freqs) or NULL to evaluate exp() on the device. */
int ddpm3d_k_timestep_embedding(const float* t, const float* freqs, float* out, int B, int dim, void* stream);

/* QKVAttentionLegacy / QKVAttention core (unet.py:328-393): qkv [B][T][3C] channels-last -> out [B][T][C].
 * 16-bit types with 64-wide heads run the fused tcgen05 kernel; new_order | 0x100 forces the CUDA-core kernel. */
int ddpm3d_k_attention(int dtype, const void* qkv, void* out, int B, int T, int C, int heads, int new_order, void* stream);
/* The same with the queries restricted to tokens [q_begin, q_begin + q_count) of the T keys (what a z-slab rank runs
 * after the K/V all-gather); out is [B][q_count][C]. */
int ddpm3d_k_attention_window(int dtype, const void* qkv, void* out, int B, int T, int C, int heads, int new_order,
                              int q_begin, int q_count, void* stream);

/* ---- the steps either side of the loop (scripts/test.py) and the ensemble reduction ----------------
 * extract_patch: scripts/test.py:205-230 -- vol device fp32 (D,H,W); out (P,P,P) in (Z,H,W) order, zero padded. */
int ddpm3d_k_extract_patch(const float* vol, int D, int H, int W, int z0, int h0, int w0, int P, float* out, void* stream);
/* hann_accumulate: scripts/test.py:91-137 -- patch (P,P,P) (Z,H,W) order; window = np.hanning(P) as device fp64 [P],
 * window_max = max(window)^3; arr / wsum device fp32 in the reference's (H,W,Z) output order.  Bit-identical to
 * the reference's numpy arithmetic (fp32 * fp64 -> fp64 -> stored fp32). */
int ddpm3d_k_hann_accumulate(const float* patch, const double* window, double window_max, int P, int D, int H, int W,
                             int z0, int h0, int w0, float* arr, float* wsum, void* stream);
/* hann_finalize: scripts/test.py:139 -- arr /= wsum where wsum > 0. */
int ddpm3d_k_hann_finalize(float* arr, const float* wsum, int64_t n, void* stream);
/* Uncertainty-map ensemble (README.md:44; BASELINE config 5): voxel-wise Welford update with one more sample
 * (count_after = number of samples including x) and the merge of two partials (a <- a (+) b). */
int ddpm3d_k_welford_update(float* mean, float* m2, const float* x, int count_after, int64_t n, void* stream);
int ddpm3d_k_welford_merge(float* mean_a, float* m2_a, int na, const float* mean_b, const float* m2_b, int nb, int64_t n,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DDPM3D_H_ */
