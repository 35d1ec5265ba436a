// Launchers of every device kernel on the path.  All tensors are channels-last (NDHWC) unless
// stated.  `dt` is DDPM3D_FP32 or DDPM3D_BF16 and selects the activation / conv-weight element.
#pragma once

#include "common.cuh"

namespace ddpm3d {

// Residual handling in the conv epilogue (ResBlock skip path, unet.py:238-256).
enum ResMode { RES_NONE = 0, RES_SAME = 1, RES_POOL = 2 /* avg 2x2 of a (2Ho,2Wo) tensor */, RES_UP = 3 /* nearest from (Ho/2,Wo/2) */ };
// Resampling of a GroupNorm/SiLU result before it is written (h_upd, unet.py:240-242).
enum Resample { RS_NONE = 0, RS_POOL = 1, RS_UP = 2 };

struct ConvSrc {
  const void* ptr = nullptr;  // [B][Z][Hin][Win][C]
  int C = 0;
};

struct ConvArgs {
  int dt = DDPM3D_FP32;       // element type of the main source and of its weight columns
  // element type of the extra (1x1x1) sources and THEIR weight columns, of the residual and of the output: -1 = dt.
  // In the default 16-bit mode the conv operands (GroupNorm outputs, weights) are bf16 while block inputs / outputs
  // -- tensors that only GroupNorm, the skip path and the residual read -- are stored as fp16.
  int dt_io = -1;
  int io_dt() const { return dt_io < 0 ? dt : dt_io; }
  ConvSrc main;               // 3x3x3 (taps=27) or 1x1x1 (taps=1) source
  int in_zpad = 0;            // main source carries this many halo planes on each side of Z (z-slab sharding):
                              // [B][Z + 2*in_zpad][Hin][Win][C]; the conv then never pads in Z itself
  int taps = 27;
  int stride_hw = 1;          // Downsample(use_conv=True): (1,2,2)  (unet.py:129-133)
  ConvSrc extra[2];           // 1x1x1 sources appended along K (skip_connection folded in; K11 concat elision)
  int n_extra = 0;
  int extra_is_identity = 0;  // the extra source is the block's identity skip folded in with unit weights (profiling only)
  const void* w = nullptr;    // [Cout][w_ld], first Ktot = taps*main.C + sum(extra.C) columns used; k = tap*C + ci
  int w_ld = 0;               // row pitch of w in elements (0 = Ktot)
  const float* bias = nullptr;  // [Cout] (already includes the folded skip bias)
  const void* residual = nullptr;  // element type dt
  int res_mode = RES_NONE;
  void* out = nullptr;        // dt, channels-last; or fp32 planar NCDHW when out_planar_f32
  int out_planar_f32 = 0;
  int B = 0, Z = 0, Ho = 0, Wo = 0, Cout = 0;  // output geometry; input H/W = Ho*stride
  // Optional (tcgen05 path only): per-channel [sum, sum of squares] of the OUTPUT, accumulated from the fp32
  // accumulators in the epilogue, one partial per CTA: chsum_out[B][chsum_slots()][Cout][2].  The consuming
  // GroupNorm then skips its statistics pass over the tensor.  chsum_written is set by the launcher.  The sums are
  // taken over (x - bias[c]), i.e. over the accumulators: pass the same `bias` to GnArgs::chsum_bias.
  float* chsum_out = nullptr;
  int chsum_written = 0;
  // split-K scratch for layers with too few tiles to fill the GPU (fp32 partial tiles); see conv_tc_scratch_bytes
  float* splitk_scratch = nullptr;
  size_t splitk_bytes = 0;
  int splitk_allowed = 0;
  int strip_allowed = 2;    // (unit-test entry point default; the engine passes its `strip` option) strip variant: A staged once per (dz, chunk),
                            // 9 in-plane taps by row-shifted descriptors.  1 = large layers only, 2 = also 12..23-wide planes (two z-planes per tile)
  int strip_maxw = 4;       // strip variant: most stages of the weight ring (as many as fit beside the two strips are used)
  int pdl = 0;              // tcgen05 kernels: launch with programmatic stream serialisation (the prologue overlaps the tail of the
                            // kernel before; every thread executes griddepcontrol.wait before it touches global memory)
  int cluster_allowed = 1;  // 2-CTA clusters sharing the weight tile by TMA multicast (big layers)
  int head_v2_allowed = 1;  // head conv: 32x16x4 bricks, 8 voxels per thread, cp.async double-buffered channel stages
  int stem_tc_allowed = 1;  // Cin == 2 stem as one M128 x Cout x K64 tcgen05 tile per 128 voxels (16-bit modes)
};
inline int chsum_slots() { return sm_count(); }  // one per persistent CTA (unused slots are zeroed by the launcher)

int conv_simt(const ConvArgs& a, cudaStream_t s);
// thin ends of the network on the CUDA cores with smem-staged halo bricks (conv_small.cu)
bool conv_stem_eligible(const ConvArgs& a);  // Cin == 2
int conv_stem(ConvArgs& a, cudaStream_t s);
int conv_stem_chsum_slots(const ConvArgs& a);  // slots of chsum_out the tensor-core stem fills (0 = none)
bool conv_head_eligible(const ConvArgs& a);  // Cout <= 2, fp32, planar output
int conv_head(const ConvArgs& a, cudaStream_t s);
// tcgen05 path; returns DDPM3D_ERR_ARG (without launching) when the shape is not eligible.
bool conv_tc_eligible(const ConvArgs& a);
int conv_tc(ConvArgs& a, cudaStream_t s);
size_t conv_tc_scratch_bytes(const ConvArgs& a);
int conv_tc_plan_query(const ConvArgs& a, int sms, int* out8);  // kernel / tiling conv_tc() would pick (host arithmetic only)
// test-only probe of row-shifted SWIZZLE_128B operand descriptors (see conv_tc.cu)
int probe_rowshift(const void* a, int rows, const void* ident, int shift, int mode, float* out, cudaStream_t s);  // split-K scratch this launch wants (0 = no split)

// ---- GroupNorm32 + FiLM + SiLU (K4/K5/K6) ------------------------------------------------------
struct GnArgs {
  int dt = DDPM3D_FP32;
  const void* src[2] = {nullptr, nullptr};  // virtual channel concat of up to two tensors (K11)
  int C[2] = {0, 0};
  int B = 0, Z = 0, H = 0, W = 0;           // input geometry
  const float* gamma = nullptr;             // [Ctot]
  const float* beta = nullptr;
  const float* film = nullptr;              // [B][film_stride] : scale at +0, shift at +Ctot (unet.py:248-252); or NULL
  int64_t film_stride = 0;
  const float* pre_add = nullptr;           // [B][pre_stride]: h + emb_out before the norm (use_scale_shift_norm=False, unet.py:253-255)
  int64_t pre_stride = 0;
  int silu = 1;
  int resample = RS_NONE;
  void* out = nullptr;                      // dt_out (or fp32 when out_f32)
  int dt_out = -1;                          // 16-bit output format when it differs from the input's (-1 = dt)
  int out_f32 = 0;
  int out_zpad = 0;                         // output tensor has this many halo planes on each side of Z (left untouched)
  int stream_allowed = 1;                   // large tensors: the TMA streaming apply kernel (gn_stream.cu)
  int stream_min_mb = 48;                   // ... from this many MB of input on
  int pdl = 1;                              // the apply kernel is launched with programmatic stream serialisation (it follows
                                            // the finalize kernel, which releases it early); 0 when another kernel sits between
  // cross-rank statistics (z-slab sharding): when gathered != NULL the finalize pass reads
  // gathered[world][B][32][2] (fp64 sums, rank order) instead of the local partials
  const double* gathered = nullptr;
  int world = 1;
  double inv_count_global = 0.0;
  // peer path: gathered is this rank's mailbox (rank stride gather_stride doubles); the finalize kernel first waits
  // until gather_flags[r] >= *gather_seq for every rank r
  int64_t gather_stride = 0, gather_parity_stride = 0;  // `gathered` holds two such slots, selected by (*gather_seq & 1)
  const uint32_t* gather_flags = nullptr;
  const uint32_t* gather_seq = nullptr;                 // device counter, incremented by the push kernel
  // peer path: base (batch 0) of the upper neighbour's trailing halo plane / the lower neighbour's leading halo plane of
  // the tensor that corresponds to `out` (same layout), or NULL
  void* peer_halo[2] = {nullptr, nullptr};
  // per-source channel sums produced by the preceding convolutions' epilogues (see ConvArgs::chsum_out); when
  // every source has them the statistics pass is skipped
  const float* chsum[2] = {nullptr, nullptr};
  int chsum_P[2] = {0, 0};                  // slots per source (0 = chsum_slots(): one per SM; the stem has one per CTA)
  // the sums are over (x - bias_c) of the producing convolution (no cancellation when a bias dominates): its bias [C_s]
  const float* chsum_bias[2] = {nullptr, nullptr};
  // scratch (owned by the caller / workspace)
  double* partials = nullptr;               // [B][32][2][n_chunks] fp64 (pivoted fp32 sums folded back in fp64)
  float* ab = nullptr;                      // [B][2][Ctot]
  int n_chunks = 0;
};
int gn_chunks(int64_t rows_per_batch);      // number of stats chunks per batch element
int gn_forward(const GnArgs& a, cudaStream_t s, int* launches);
// split form for the sharded path: stats -> local fp64 sums [B][32][2] -> (all-gather) -> finalize + apply
int gn_stats_local(const GnArgs& a, double* sums, cudaStream_t s);
int gn_chsum_local(const GnArgs& a, double* sums, cudaStream_t s);  // same, from the producers' channel sums
int gn_finalize_apply(const GnArgs& a, cudaStream_t s);
// statistics from the producers' channel sums: finalize + apply only (one read + one write of the tensor)
int gn_forward_chsum(const GnArgs& a, cudaStream_t s);
// apply pass of the large tensors as a TMA streaming kernel (16-bit in / out, no resampling, channel counts % 128 == 0)
bool gn_apply_stream_eligible(const GnArgs& a);
int gn_apply_stream(const GnArgs& a, cudaStream_t s);
// statistics (from the channel sums when present, else a pass over the tensor) + the per-(b, c) affine a.ab, no apply
int gn_finalize_only(const GnArgs& a, cudaStream_t s);

// ---- the fused head (head_tc.cu): GroupNorm apply -> SiLU -> Conv3d(C -> 1|2) on tcgen05, 16-bit modes ------------
// x: the last block output [B][Z][H][W][C] (dt_src); ab: [B][2][C] from gn_finalize_only; w: fp32 [Cout][27*C];
// out: planar fp32 (B, Cout, Z, H, W)
bool conv_head_tc_eligible(int dt_src, int C, int Cout);
int conv_head_tc(int dt_src, const void* x, const float* ab, const float* w, const float* bias, float* out, int B, int Z, int H,
                 int W, int C, int Cout, cudaStream_t s);

// plain resample (Upsample(use_conv=True) front half, unet.py:100-105)
int resample_hw(int dt, const void* in, void* out, int B, int Z, int H, int W, int C, int mode, cudaStream_t s);

// ---- network input / embedding -----------------------------------------------------------------
// cat([x, low_res], 1) + cast (unet.py:1690-1693,1035): two fp32 (B,1,Z,H,W) -> [B][Z][H][W][2]
// out_zpad: halo planes on each side of Z in `out` ([B][Z+2p][H][W][2]); plane = H*W voxels
int pack_input(int dt, const float* x, const float* low, void* out, int B, int Z, int64_t plane, int out_zpad, cudaStream_t s,
               void* peer_lo = nullptr, void* peer_hi = nullptr);
int pack_input_planar(int dt, const float* x, const float* low, int Cx, void* out, int B, int Z, int64_t plane, int out_zpad,
                      cudaStream_t s);

struct EmbArgs {
  const float* t = nullptr;          // [B]
  const float* freqs = nullptr;      // [model_channels/2] host-computed frequencies, or NULL
  const int64_t* y = nullptr;        // [B] or NULL
  int B = 0, model_channels = 0, ted = 0;
  const float *w0, *b0, *w2, *b2;    // time_embed.{0,2}
  const float* label_emb = nullptr;  // [num_classes][ted]
  float* emb_silu = nullptr;         // [B][ted]  = SiLU(emb)   (every consumer applies SiLU first, unet.py:199-205)
  const float* w_all = nullptr;      // all emb_layers.1.weight stacked: [rows_total][ted]
  const float* b_all = nullptr;      // [rows_total]
  int rows_total = 0;
  float* emb_out = nullptr;          // [B][rows_total]
};
int timestep_embedding_k(const float* t, const float* freqs, float* out, int B, int dim, cudaStream_t s);
int embedding_forward(const EmbArgs& a, cudaStream_t s, int* launches);

// ---- sampler update (K9) ---------------------------------------------------------------------
struct UpdateArgs {
  const float* x = nullptr;
  const float* model_out = nullptr;
  const float* noise = nullptr;          // or NULL with use_philox
  const int32_t* t_index = nullptr;      // device [B]  (or NULL: use *step_counter for every b)
  const int32_t* step_counter = nullptr; // device scalar
  const ddpm3d_step_scalars* table = nullptr;  // device [T]
  int mean_type = DDPM3D_MEAN_EPSILON, var_type = DDPM3D_VAR_LEARNED_RANGE, clip = 1;
  float* sample = nullptr; float* pred_xstart = nullptr; float* mean = nullptr; float* log_variance = nullptr;
  int B = 0, C = 1; int64_t n = 0;       // n = spatial size per (b,c)
  int64_t noise_step_stride = 0;         // with step_counter: noise += exec_index * stride
  int T = 0;
  int ddim = 0; float eta = 0.f;         // gaussian_diffusion.py:537-585 instead of :430-438
  int use_philox = 0; uint64_t seed = 0;
  // Philox counter = global element index b*idx_bstride + idx_offset + (local offset): slabs of one volume draw
  // disjoint parts of one noise field.  idx_bstride == 0 -> the local index (single GPU).
  int64_t idx_offset = 0, idx_bstride = 0;
};
int p_sample_update_k(const UpdateArgs& a, cudaStream_t s);
int step_set_k(int32_t* step_counter, float* t_model, const ddpm3d_step_scalars* table, int B, int index, int exec, cudaStream_t s);
int step_from_tensor_k(const int64_t* t, int32_t* t_index, float* t_model, const ddpm3d_step_scalars* table, int B, int T,
                       cudaStream_t s);
int step_advance_k(int32_t* step_counter, float* t_model, const ddpm3d_step_scalars* table, int B, cudaStream_t s);

// ---- attention core (K12) --------------------------------------------------------------------
// scratch (optional): attention_tc_scratch_bytes() bytes for the tensor-core path (64-wide heads, 16-bit types);
// without it, or for other head widths / fp32, the CUDA-core kernel runs
// q_begin / q_count: the queries are tokens [q_begin, q_begin + q_count) of the T_tok keys (z-slab sharding: local
// queries against the all-gathered keys / values); out is [B][q_count][C].  q_count < 0 = all tokens.
int attention_k(int dt, const void* qkv, void* out, int B, int T_tok, int C, int heads, int new_order, void* scratch,
                size_t scratch_bytes, cudaStream_t s, int q_begin = 0, int q_count = -1);
size_t attention_tc_scratch_bytes(int dt, int B, int T, int C, int heads);
int attention_tc(int dt, const void* qkv, void* out, int B, int T, int C, int heads, int new_order, void* scratch, cudaStream_t s,
                 int q_begin = 0, int q_count = -1);

}  // namespace ddpm3d
