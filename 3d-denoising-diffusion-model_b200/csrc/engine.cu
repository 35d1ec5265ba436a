// The ddpm3d context: native topology builder (restating UNetModel_noatt.__init__,
// guided_diffusion/unet.py:751-997), state_dict ingestion and weight packing, the activation
// workspace, one UNet evaluation (unet.py:1015-1044,1687-1694) as a sequence of fused kernels,
// and the device-resident sampler (gaussian_diffusion.py:395-535).  Also the C ABI (include/ddpm3d.h).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "comm.h"
#include "kernels.h"

namespace ddpm3d {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }

namespace {

// ------------------------------------------------------------------------------------------------
// topology
// ------------------------------------------------------------------------------------------------
enum LayerKind { L_CONV = 0, L_RES = 1, L_ATTN = 2, L_UPCONV = 3 };

struct DevConv {
  void* w = nullptr;      // dt [Cout][Ktot]
  float* bias = nullptr;  // [Cout]
  int Ktot = 0;
  int taps = 27;          // 27 (Conv3d 3x3x3), 9 (Conv2d 3x3 of a dims = 2 network) or 1
};

struct Layer {
  int kind = L_CONV;
  std::string prefix;
  int cin = 0, cout = 0;
  bool up = false, down = false;
  int stride_hw = 1;
  int heads = 0;
  // device weights (after finalize)
  float *gn1_g = nullptr, *gn1_b = nullptr, *gn2_g = nullptr, *gn2_b = nullptr;
  DevConv c1, c2;      // res: in_layers.2 / out_layers.3 (+ folded skip); conv/upconv: c1; attn: c1 = qkv, c2 = proj_out
  bool skip_conv = false;
  int emb_off = 0, emb_rows = 0;  // rows of this block in the stacked emb_layers matrix
};

struct Param {
  std::string key;
  std::vector<int64_t> shape;
  std::vector<float> host;
  bool loaded = false;
  int64_t numel() const {
    int64_t n = 1;
    for (auto d : shape) n *= d;
    return n;
  }
};

struct Act {
  void* p = nullptr;
  int C = 0, H = 0, W = 0;
  float* chsum = nullptr;  // per-CTA channel sums of this tensor from the producing conv's epilogue (or NULL)
  int cs_slots = 0;        // slots of chsum (0 = chsum_slots())
  const float* cs_bias = nullptr;  // that conv's bias: the sums are over (x - bias)
};

struct Arena {
  char* base = nullptr;
  size_t off = 0, cap = 0, peak = 0;
  bool dry = true;
  bool overflow = false;
  void* alloc(size_t bytes) {
    bytes = (bytes + 255) & ~size_t(255);
    size_t o = off;
    off += bytes;
    peak = std::max(peak, off);
    if (dry) return nullptr;
    if (off > cap) { overflow = true; return nullptr; }
    return base + o;
  }
};

uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // NaN
  const uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (uint16_t)(u >> 16);
}

uint16_t f32_to_f16_rn(float f) {
  uint32_t x;
  std::memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  const uint32_t mant = x & 0x007fffffu;
  const int32_t exp = (int32_t)((x >> 23) & 0xff) - 127 + 15;
  if (((x >> 23) & 0xff) == 0xff) return (uint16_t)(sign | 0x7c00u | (mant ? 0x200u : 0));  // inf / nan
  if (exp >= 31) return (uint16_t)(sign | 0x7c00u);                                           // overflow -> inf
  if (exp <= 0) {                                                                              // subnormal / zero
    if (exp < -10) return (uint16_t)sign;
    const uint32_t m = mant | 0x00800000u;
    const int shift = 14 - exp;
    uint32_t h = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1))) ++h;
    return (uint16_t)(sign | h);
  }
  uint32_t h = ((uint32_t)exp << 10) | (mant >> 13);
  const uint32_t rem = mant & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) ++h;  // may carry into the exponent: still correct
  return (uint16_t)(sign | h);
}

struct GraphKey {
  int kind;  // 0 forward, 1 p_sample step, 2 loop step
  int B, Z, H, W;
  const void* p[8];
  int64_t i[3];
  bool operator<(const GraphKey& o) const { return std::memcmp(this, &o, sizeof(GraphKey)) < 0; }
};

struct ProfEntry {
  cudaEvent_t a, b;
  int kind;
  double work;
};

}  // namespace
}  // namespace ddpm3d

using namespace ddpm3d;

struct ddpm3d_ctx {
  ddpm3d_config cfg{};
  int stem_cin() const { return cfg.in_channels * (cfg.unconditional ? 1 : 2); }
  bool two_d() const { return cfg.dims == 2; }
  // conv weight shape in the state_dict: [Cout][Cin][k][k][k] (dims 3) or [Cout][Cin][k][k] (dims 2)
  std::vector<int64_t> wshape(int64_t co, int64_t ci, int64_t k) const {
    return two_d() ? std::vector<int64_t>{co, ci, k, k} : std::vector<int64_t>{co, ci, k, k, k};
  }
  std::vector<Param> params;
  std::unordered_map<std::string, int> index;
  std::vector<std::vector<Layer>> input_blocks, output_blocks;
  std::vector<Layer> middle;
  int out_norm_ch = 0, out_conv_in = 0, ted = 0;
  bool finalized = false;
  int device = -1;
  int dt = DDPM3D_FP32;   // conv operands: GroupNorm outputs and the 3x3x3 / qkv / proj weights
  int dts = DDPM3D_FP32;  // block inputs / outputs and the tensor between the two convs of a ResBlock (read only by GroupNorm,
                          // the skip path and residual adds): fp16 in the default bf16 mode -- same bytes, 3 more mantissa bits
  size_t esz = 4;
  std::vector<void*> dev_allocs;
  // embedding path
  float *te_w0 = nullptr, *te_b0 = nullptr, *te_w2 = nullptr, *te_b2 = nullptr, *label_emb = nullptr;
  float *emb_w_all = nullptr, *emb_b_all = nullptr;
  int emb_rows_total = 0;
  float *out_gn_g = nullptr, *out_gn_b = nullptr;
  DevConv out_conv;  // out.2 (fp32 always)
  // workspace
  char* ws = nullptr;
  size_t ws_cap = 0;
  // sampler
  ddpm3d_step_scalars* d_table = nullptr;
  int T = 0, mean_type = DDPM3D_MEAN_EPSILON, var_type = DDPM3D_VAR_LEARNED_RANGE;
  int ddim = 0;
  float eta = 0.f;
  float* d_tmodel = nullptr;
  int32_t* d_tindex = nullptr;
  int32_t* d_counter = nullptr;
  int loop_B = 0;
  float* d_mo = nullptr;
  size_t mo_cap = 0;
  float* d_img = nullptr;
  size_t img_cap = 0;
  // options
  int use_graph = 1, conv_path = 0, profile = 0, fuse_stats = 1, split_k = 1, cluster = 0, strip = 2, strip_w = 4, fold_identity = 1, stem_tc = 1, head_v2 = 1,
      head_tc = 1, slab_p2p = 1, pdl = 2, gn_stream = 1, gn_stream_mb = 48;
  float* splitk_buf = nullptr;  // fp32 partial tiles of split-K convolutions (sized by the dry run)
  size_t splitk_cap = 0, splitk_need = 0;
  cudaStream_t cap_stream = nullptr;
  float* d_freqs = nullptr;  // timestep_embedding frequencies computed by the host exactly like nn.py:113-115
  int n_freqs = 0;
  int64_t launches = 0;
  SlabComm slab;  // z-slab sharding of one volume over ranks (inactive unless ddpm3d_set_comm was called with world > 1)
  std::map<GraphKey, cudaGraphExec_t> graphs;
  std::map<GraphKey, int64_t> graph_launches;
  std::vector<ProfEntry> prof;
};

namespace ddpm3d {
namespace {

// ------------------------------------------------------------------------------------------------
// constructor replay (unet.py:797-997)
// ------------------------------------------------------------------------------------------------
int heads_for(const ddpm3d_config& c, int ch, int n) {
  if (c.num_head_channels == -1) return n;
  return ch / c.num_head_channels;
}

Layer make_res(const std::string& prefix, int cin, int cout, bool up = false, bool down = false) {
  Layer L;
  L.kind = L_RES; L.prefix = prefix; L.cin = cin; L.cout = cout; L.up = up; L.down = down;
  L.skip_conv = cin != cout;
  return L;
}
Layer make_attn(const std::string& prefix, int ch, int heads) {
  Layer L;
  L.kind = L_ATTN; L.prefix = prefix; L.cin = ch; L.cout = ch; L.heads = heads;
  return L;
}
Layer make_conv(int kind, const std::string& prefix, int cin, int cout, int stride) {
  Layer L;
  L.kind = kind; L.prefix = prefix; L.cin = cin; L.cout = cout; L.stride_hw = stride;
  return L;
}

bool is_attention_ds(const ddpm3d_config& c, int ds) {
  for (int i = 0; i < c.n_attention_ds; ++i)
    if (c.attention_ds[i] == ds) return true;
  return false;
}

void add_param(ddpm3d_ctx* ctx, const std::string& key, std::vector<int64_t> shape) {
  Param p;
  p.key = key;
  p.shape = std::move(shape);
  ctx->index[key] = (int)ctx->params.size();
  ctx->params.push_back(std::move(p));
}

void add_layer_params(ddpm3d_ctx* ctx, const Layer& L) {
  const std::string& p = L.prefix;
  const int64_t ted = ctx->ted;
  if (L.kind == L_CONV || L.kind == L_UPCONV) {
    add_param(ctx, p + ".weight", ctx->wshape(L.cout, L.cin, 3));
    add_param(ctx, p + ".bias", {L.cout});
  } else if (L.kind == L_RES) {
    const int64_t e = ctx->cfg.use_scale_shift_norm ? 2 * L.cout : L.cout;
    add_param(ctx, p + ".in_layers.0.weight", {L.cin});
    add_param(ctx, p + ".in_layers.0.bias", {L.cin});
    add_param(ctx, p + ".in_layers.2.weight", ctx->wshape(L.cout, L.cin, 3));
    add_param(ctx, p + ".in_layers.2.bias", {L.cout});
    add_param(ctx, p + ".emb_layers.1.weight", {e, ted});
    add_param(ctx, p + ".emb_layers.1.bias", {e});
    add_param(ctx, p + ".out_layers.0.weight", {L.cout});
    add_param(ctx, p + ".out_layers.0.bias", {L.cout});
    add_param(ctx, p + ".out_layers.3.weight", ctx->wshape(L.cout, L.cout, 3));
    add_param(ctx, p + ".out_layers.3.bias", {L.cout});
    if (L.skip_conv) {
      add_param(ctx, p + ".skip_connection.weight", ctx->wshape(L.cout, L.cin, 1));
      add_param(ctx, p + ".skip_connection.bias", {L.cout});
    }
  } else {  // attention (unet.py:259-305)
    add_param(ctx, p + ".norm.weight", {L.cin});
    add_param(ctx, p + ".norm.bias", {L.cin});
    add_param(ctx, p + ".qkv.weight", {3 * L.cin, L.cin, 1});
    add_param(ctx, p + ".qkv.bias", {3 * L.cin});
    add_param(ctx, p + ".proj_out.weight", {L.cin, L.cin, 1});
    add_param(ctx, p + ".proj_out.bias", {L.cin});
  }
}

int build_topology(ddpm3d_ctx* ctx) {
  const ddpm3d_config& c = ctx->cfg;
  DD_CHECK(c.n_levels >= 1 && c.n_levels <= DDPM3D_MAX_LEVELS, DDPM3D_ERR_ARG, "config: n_levels out of range");
  DD_CHECK(c.model_channels > 0 && c.model_channels % 32 == 0, DDPM3D_ERR_ARG,
           "config: model_channels must be a positive multiple of 32 (GroupNorm32)");
  DD_CHECK(c.num_res_blocks >= 1, DDPM3D_ERR_ARG, "config: num_res_blocks must be >= 1");
  DD_CHECK(c.in_channels >= 1 && c.in_channels <= 64, DDPM3D_ERR_ARG, "config: in_channels out of range");
  DD_CHECK(c.dims == 0 || c.dims == 2 || c.dims == 3, DDPM3D_ERR_ARG, "config: dims must be 2 or 3");
  DD_CHECK(c.out_channels >= 1, DDPM3D_ERR_ARG, "config: out_channels must be >= 1");
  const int mc = c.model_channels;
  ctx->ted = mc * 4;
  const int heads_up = c.num_heads_upsample == -1 ? c.num_heads : c.num_heads_upsample;
  int ch = c.channel_mult[0] * mc;
  const int input_ch = ch;
  ctx->input_blocks.push_back({make_conv(L_CONV, "input_blocks.0.0", ctx->stem_cin(), ch, 1)});
  std::vector<int> chans{ch};
  int ds = 1;
  for (int level = 0; level < c.n_levels; ++level) {
    for (int r = 0; r < c.num_res_blocks; ++r) {
      const int n = (int)ctx->input_blocks.size();
      const int cout = c.channel_mult[level] * mc;
      std::vector<Layer> blk{make_res("input_blocks." + std::to_string(n) + ".0", ch, cout)};
      ch = cout;
      if (is_attention_ds(c, ds)) {
        DD_CHECK(c.num_head_channels == -1 || ch % c.num_head_channels == 0, DDPM3D_ERR_ARG,
                 "config: channels not divisible by num_head_channels");
        blk.push_back(make_attn("input_blocks." + std::to_string(n) + ".1", ch, heads_for(c, ch, c.num_heads)));
      }
      ctx->input_blocks.push_back(blk);
      chans.push_back(ch);
    }
    if (level != c.n_levels - 1) {
      const int n = (int)ctx->input_blocks.size();
      if (c.resblock_updown)
        ctx->input_blocks.push_back({make_res("input_blocks." + std::to_string(n) + ".0", ch, ch, false, true)});
      else
        ctx->input_blocks.push_back({make_conv(L_CONV, "input_blocks." + std::to_string(n) + ".0.op", ch, ch, 2)});
      chans.push_back(ch);
      ds *= 2;
    }
  }
  if (c.middle_attention) {  // UNetModel (unet.py:539-563): ResBlock, AttentionBlock, ResBlock
    DD_CHECK(c.num_head_channels == -1 || ch % c.num_head_channels == 0, DDPM3D_ERR_ARG,
             "config: channels not divisible by num_head_channels");
    ctx->middle = {make_res("middle_block.0", ch, ch), make_attn("middle_block.1", ch, heads_for(c, ch, c.num_heads)),
                   make_res("middle_block.2", ch, ch)};
  } else {
    ctx->middle = {make_res("middle_block.0", ch, ch), make_res("middle_block.1", ch, ch)};
  }
  int outch = ch;
  for (int level = c.n_levels - 1; level >= 0; --level) {
    for (int i = 0; i < c.num_res_blocks + 1; ++i) {
      const int inch = chans.back();
      chans.pop_back();
      if (!chans.empty()) { outch = chans.back(); chans.pop_back(); } else outch = inch;
      const int n = (int)ctx->output_blocks.size();
      const std::string pre = "output_blocks." + std::to_string(n) + ".";
      std::vector<Layer> blk{make_res(pre + "0", inch * 2, outch)};
      if (is_attention_ds(c, ds)) blk.push_back(make_attn(pre + std::to_string(blk.size()), outch, heads_for(c, outch, heads_up)));
      if (level && i == c.num_res_blocks) {
        const std::string k = std::to_string(blk.size());
        if (c.resblock_updown) blk.push_back(make_res(pre + k, outch, outch, true, false));
        else blk.push_back(make_conv(L_UPCONV, pre + k + ".conv", outch, outch, 1));
        ds /= 2;
      }
      ctx->output_blocks.push_back(blk);
      chans.push_back(outch);
    }
  }
  ctx->out_norm_ch = outch;
  ctx->out_conv_in = input_ch;
  DD_CHECK(ctx->out_norm_ch == ctx->out_conv_in, DDPM3D_ERR_ARG, "config: out norm / conv channel mismatch (unet.py:993-997)");

  // parameters in the reference's registration order
  const int64_t ted = ctx->ted;
  add_param(ctx, "time_embed.0.weight", {ted, mc});
  add_param(ctx, "time_embed.0.bias", {ted});
  add_param(ctx, "time_embed.2.weight", {ted, ted});
  add_param(ctx, "time_embed.2.bias", {ted});
  if (c.num_classes > 0) add_param(ctx, "label_emb.weight", {c.num_classes, ted});
  for (auto& blk : ctx->input_blocks)
    for (auto& L : blk) add_layer_params(ctx, L);
  for (auto& L : ctx->middle) add_layer_params(ctx, L);
  for (auto& blk : ctx->output_blocks)
    for (auto& L : blk) add_layer_params(ctx, L);
  add_param(ctx, "out.0.weight", {ctx->out_norm_ch});
  add_param(ctx, "out.0.bias", {ctx->out_norm_ch});
  add_param(ctx, "out.2.weight", ctx->wshape(c.out_channels, ctx->out_conv_in, 3));
  add_param(ctx, "out.2.bias", {c.out_channels});
  return DDPM3D_OK;
}

// ------------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------------
const Param* find(ddpm3d_ctx* ctx, const std::string& key) {
  auto it = ctx->index.find(key);
  return it == ctx->index.end() ? nullptr : &ctx->params[it->second];
}

int upload(ddpm3d_ctx* ctx, const void* host, size_t bytes, void** out) {
  void* d = nullptr;
  DD_CUDA(cudaMalloc(&d, std::max<size_t>(bytes, 16)));
  ctx->dev_allocs.push_back(d);
  DD_CUDA(cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));
  *out = d;
  return DDPM3D_OK;
}

int upload_f32(ddpm3d_ctx* ctx, const std::string& key, float** out) {
  const Param* p = find(ctx, key);
  DD_CHECK(p && p->loaded, DDPM3D_ERR_MISSING, "missing state_dict tensor: " + key);
  return upload(ctx, p->host.data(), p->host.size() * sizeof(float), (void**)out);
}

// [Cout][Cin][taps] (reference) -> [Cout][taps*Cin (+ Cskip)], element type dt; bias (+ skip bias) fp32
int pack_conv(ddpm3d_ctx* ctx, int dt, const std::string& wkey, const std::string& bkey, int taps,
              const std::string* skip_w, const std::string* skip_b, DevConv* out, bool append_identity = false,
              int dt_extra = -1) {
  if (dt_extra < 0) dt_extra = dt;
  const Param* w = find(ctx, wkey);
  const Param* b = find(ctx, bkey);
  DD_CHECK(w && w->loaded, DDPM3D_ERR_MISSING, "missing state_dict tensor: " + wkey);
  DD_CHECK(b && b->loaded, DDPM3D_ERR_MISSING, "missing state_dict tensor: " + bkey);
  const int64_t Cout = w->shape[0], Cin = w->shape[1];
  // dims = 2 (unet.py:396-716 with Conv2d): the images run through the same kernels as (B, C, 1, H, W) volumes, with
  // the 9 in-plane taps only (k = tap9 * Cin + ci)
  if (taps == 27 && w->shape.size() == 4) taps = 9;
  out->taps = taps;
  const Param* sw = nullptr;
  const Param* sb = nullptr;
  int64_t Cs = 0;
  if (skip_w) {
    sw = find(ctx, *skip_w);
    sb = find(ctx, *skip_b);
    DD_CHECK(sw && sw->loaded, DDPM3D_ERR_MISSING, "missing state_dict tensor: " + *skip_w);
    DD_CHECK(sb && sb->loaded, DDPM3D_ERR_MISSING, "missing state_dict tensor: " + *skip_b);
    Cs = sw->shape[1];
  }
  // append_identity: skip_connection = Identity written as a 1x1x1 conv with unit weights, so that the residual
  // can be folded into the accumulation like a real skip conv (x * 1.0 is exact in the fp32 accumulator)
  if (append_identity) Cs = Cout;
  const int64_t Ktot = taps * Cin + Cs;
  std::vector<float> packed((size_t)(Cout * Ktot), 0.f);
  for (int64_t co = 0; co < Cout; ++co) {
    float* row = packed.data() + co * Ktot;
    {
      const float* src = w->host.data() + co * Cin * taps;
      for (int64_t ci = 0; ci < Cin; ++ci)
        for (int t = 0; t < taps; ++t) row[(int64_t)t * Cin + ci] = src[ci * taps + t];
    }
    if (sw)
      for (int64_t ci = 0; ci < Cs; ++ci) row[taps * Cin + ci] = sw->host[co * Cs + ci];
    else if (append_identity)
      row[taps * Cin + co] = 1.0f;
  }
  std::vector<float> bias(b->host);
  if (sb)
    for (int64_t co = 0; co < Cout; ++co) bias[co] += sb->host[co];
  out->Ktot = (int)Ktot;
  DD_TRY(upload(ctx, bias.data(), bias.size() * sizeof(float), (void**)&out->bias));
  if (is_half_dt(dt)) {
    // the columns of the folded skip (1x1x1 over the block input) carry the block input's 16-bit format
    std::vector<uint16_t> h(packed.size());
    const int64_t kmain = taps * Cin;
    for (int64_t co = 0; co < Cout; ++co)
      for (int64_t k = 0; k < Ktot; ++k) {
        const size_t i = (size_t)(co * Ktot + k);
        h[i] = (k < kmain ? dt : dt_extra) == DDPM3D_BF16 ? f32_to_bf16_rn(packed[i]) : f32_to_f16_rn(packed[i]);
      }
    DD_TRY(upload(ctx, h.data(), h.size() * 2, &out->w));
  } else {
    DD_TRY(upload(ctx, packed.data(), packed.size() * 4, &out->w));
  }
  return DDPM3D_OK;
}

int finalize_layer(ddpm3d_ctx* ctx, Layer& L, std::vector<float>& emb_w, std::vector<float>& emb_b) {
  const std::string& p = L.prefix;
  const int dt = ctx->dt, dts = ctx->dts;
  if (L.kind == L_CONV || L.kind == L_UPCONV) {
    // stem / Downsample / Upsample convs read a block input directly: the whole conv runs in the storage format
    DD_TRY(pack_conv(ctx, dts, p + ".weight", p + ".bias", 27, nullptr, nullptr, &L.c1));
  } else if (L.kind == L_RES) {
    DD_TRY(upload_f32(ctx, p + ".in_layers.0.weight", &L.gn1_g));
    DD_TRY(upload_f32(ctx, p + ".in_layers.0.bias", &L.gn1_b));
    DD_TRY(upload_f32(ctx, p + ".out_layers.0.weight", &L.gn2_g));
    DD_TRY(upload_f32(ctx, p + ".out_layers.0.bias", &L.gn2_b));
    DD_TRY(pack_conv(ctx, dt, p + ".in_layers.2.weight", p + ".in_layers.2.bias", 27, nullptr, nullptr, &L.c1));
    if (L.skip_conv) {
      const std::string sw = p + ".skip_connection.weight", sb = p + ".skip_connection.bias";
      DD_TRY(pack_conv(ctx, dt, p + ".out_layers.3.weight", p + ".out_layers.3.bias", 27, &sw, &sb, &L.c2, false, dts));
    } else {
      // identity skip: the unit block is appended (w_ld grows by Cout); it is only read when the residual is
      // folded into the accumulation (fold_identity), otherwise the epilogue adds x and the block is skipped
      const bool ident = !L.up && !L.down && L.cin == L.cout;
      DD_TRY(pack_conv(ctx, dt, p + ".out_layers.3.weight", p + ".out_layers.3.bias", 27, nullptr, nullptr, &L.c2, ident, dts));
    }
    const Param* ew = find(ctx, p + ".emb_layers.1.weight");
    const Param* eb = find(ctx, p + ".emb_layers.1.bias");
    DD_CHECK(ew && ew->loaded && eb && eb->loaded, DDPM3D_ERR_MISSING, "missing state_dict tensor: " + p + ".emb_layers.1.*");
    L.emb_off = (int)emb_b.size();
    L.emb_rows = (int)eb->host.size();
    emb_w.insert(emb_w.end(), ew->host.begin(), ew->host.end());
    emb_b.insert(emb_b.end(), eb->host.begin(), eb->host.end());
  } else {
    DD_TRY(upload_f32(ctx, p + ".norm.weight", &L.gn1_g));
    DD_TRY(upload_f32(ctx, p + ".norm.bias", &L.gn1_b));
    DD_TRY(pack_conv(ctx, dt, p + ".qkv.weight", p + ".qkv.bias", 1, nullptr, nullptr, &L.c1));
    DD_TRY(pack_conv(ctx, dt, p + ".proj_out.weight", p + ".proj_out.bias", 1, nullptr, nullptr, &L.c2));
  }
  return DDPM3D_OK;
}

// ------------------------------------------------------------------------------------------------
// one evaluation
// ------------------------------------------------------------------------------------------------
struct Run {
  ddpm3d_ctx* ctx;
  int B, Z;
  cudaStream_t s;
  Arena arena;
  int launches = 0;
  float* emb_out = nullptr;  // [B][rows_total]

  int zp = 0;                // halo planes on conv-input tensors (1 in z-slab mode)

  size_t act_bytes(int H, int W, int C) const { return (size_t)B * Z * H * W * C * ctx->esz; }
  // a tensor a 3x3x3 conv will read: carries the halo planes in z-slab mode
  size_t conv_in_bytes(int H, int W, int C, size_t esz) const { return (size_t)B * (Z + 2 * zp) * H * W * C * esz; }

  // z-slab sharding over peer-mapped memory (comm.cu "peer path"): equal slabs, mailboxes and both neighbours' workspaces
  // mapped.  The producing kernel then stores its boundary planes into the neighbours' halo planes itself.
  bool p2p() const {
    const SlabComm& c = ctx->slab;
    return zp && ctx->slab_p2p && c.halo_p2p() && c.z_total == c.world * Z && c.z_begin == c.rank * Z;
  }
  // the same tensor in the workspace of the upper (dir 0) / lower (dir 1) neighbour (identical arena layout)
  char* peer_ptr(int dir, const void* t) const {
    const SlabComm& c = ctx->slab;
    const int nb = dir == 0 ? c.rank - 1 : c.rank + 1;
    if (nb < 0 || nb >= c.world) return nullptr;
    return c.peer_ws[dir] + ((const char*)t - ctx->ws);
  }
  // before a kernel that stores into the neighbours' halo planes of tensor `t` ([B][Z+2][plane_bytes]): handshake
  // (everything up to the previous exchange has been consumed everywhere), zero the planes at the volume's ends
  int halo_begin(void* t, size_t plane_bytes, void** peer_lo, void** peer_hi) {
    SlabComm& c = ctx->slab;
    launches += 1;
    *peer_lo = *peer_hi = nullptr;
    if (arena.dry) return DDPM3D_OK;
    prof_begin(10, 0.0);
    int r = comm_halo_pre(c, s);
    prof_end();
    DD_TRY(r);
    const size_t bstride = (size_t)(Z + 2) * plane_bytes;
    for (int b = 0; b < B; ++b) {
      if (c.rank == 0) DD_CUDA(cudaMemsetAsync((char*)t + b * bstride, 0, plane_bytes, s));
      if (c.rank + 1 == c.world) DD_CUDA(cudaMemsetAsync((char*)t + b * bstride + (size_t)(Z + 1) * plane_bytes, 0, plane_bytes, s));
    }
    if (char* up = peer_ptr(0, t)) *peer_lo = up + (size_t)(Z + 1) * plane_bytes;  // its trailing halo plane
    if (char* dn = peer_ptr(1, t)) *peer_hi = dn;                                  // its leading halo plane
    return DDPM3D_OK;
  }
  int halo_end(double bytes) {
    launches += 1;
    if (arena.dry) return DDPM3D_OK;
    prof_begin(10, bytes);
    const int r = comm_halo_post(ctx->slab, s);
    prof_end();
    return r;
  }

  // after a conv-input tensor has been produced: fetch the neighbours' boundary planes (NCCL path; on the peer path the
  // producing kernel has already stored them)
  int halo(void* t, int H, int W, int C, size_t esz) {
    if (!zp || p2p()) return DDPM3D_OK;
    return halo_nccl(t, H, W, C, esz);
  }
  int halo_nccl(void* t, int H, int W, int C, size_t esz) {
    launches += 1;
    if (arena.dry) return DDPM3D_OK;
    prof_begin(10, (double)B * 2 * H * W * C * esz);
    const int r = comm_halo_exchange(ctx->slab, t, B, Z, (size_t)H * W * C * esz, s);
    prof_end();
    return r;
  }

  void prof_begin(int kind, double work) {
    if (!ctx->profile || arena.dry) return;
    ProfEntry e;
    cudaEventCreate(&e.a);
    cudaEventCreate(&e.b);
    e.kind = kind;
    e.work = work;
    prof_record(e.a);
    ctx->prof.push_back(e);
  }
  void prof_end() {
    if (!ctx->profile || arena.dry) return;
    prof_record(ctx->prof.back().b);
  }
  // profile = 2: the step is captured like any other and the brackets become event-record NODES of the graph, so the
  // times are those of the replayed graph (no eager launch gaps between the short kernels)
  void prof_record(cudaEvent_t e) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &st);
    if (st == cudaStreamCaptureStatusActive) cudaEventRecordWithFlags(e, s, cudaEventRecordExternal);
    else cudaEventRecord(e, s);
  }

  float* alloc_chsum(int C) { return (float*)arena.alloc((size_t)B * chsum_slots() * C * 2 * sizeof(float)); }

  int conv(ConvArgs& a) {
    a.B = B;
    a.Z = Z;
    if (!ctx->fuse_stats) a.chsum_out = nullptr;
    if (a.taps != 1) a.in_zpad = zp;
    a.splitk_allowed = ctx->split_k;
    a.cluster_allowed = ctx->cluster;
    a.strip_allowed = ctx->strip;
    a.strip_maxw = ctx->strip_w;
    a.pdl = ctx->pdl >= 2 && !ctx->profile;  // (an operator bracket between two kernels would time the overlap, not the operator)
    a.stem_tc_allowed = ctx->stem_tc;
    a.head_v2_allowed = ctx->head_v2;
    ++launches;
    if (arena.dry) {
      if (is_half_dt(a.dt) && ctx->conv_path != 1 && a.splitk_allowed)
        ctx->splitk_need = std::max(ctx->splitk_need, conv_tc_scratch_bytes(a));
      return DDPM3D_OK;
    }
    a.splitk_scratch = ctx->splitk_buf;
    a.splitk_bytes = ctx->splitk_cap;
    // algorithmic flops: the unit-weight K block of a folded identity skip is an addition, not part of the contraction
    double K = (double)a.taps * a.main.C;
    for (int e = 0; e < a.n_extra; ++e) K += a.extra_is_identity ? 0.0 : a.extra[e].C;
    const double flops = 2.0 * B * Z * a.Ho * a.Wo * (double)a.Cout * K;
    const bool tc = is_half_dt(a.dt) && ctx->conv_path != 1 && conv_tc_eligible(a);
    const bool stem = !tc && ctx->conv_path != 1 && conv_stem_eligible(a);
    const bool head = !tc && ctx->conv_path != 1 && conv_head_eligible(a);
    prof_begin(tc ? 0 : ((stem || head) ? 9 : 1), flops);
    a.chsum_written = 0;
    const int r = tc ? conv_tc(a, s) : (stem ? conv_stem(a, s) : (head ? conv_head(a, s) : conv_simt(a, s)));
    prof_end();
    return r;
  }

  // statistics + affine only (g.ab); no apply pass
  int gn_stats(GnArgs& g) {
    g.B = B;
    g.Z = Z;
    g.pdl = (ctx->pdl >= 2 && ctx->profile) ? 1 : ctx->pdl;
    const int Ctot = g.C[0] + g.C[1];
    g.n_chunks = gn_chunks((int64_t)Z * g.H * g.W);
    g.partials = (double*)arena.alloc((size_t)B * g.n_chunks * 64 * sizeof(double));
    g.ab = (float*)arena.alloc((size_t)B * 2 * Ctot * sizeof(float));
    launches += (!g.pre_add && g.chsum[0] && (g.C[1] == 0 || g.chsum[1])) ? 1 : 2;  // (statistics,) finalize
    if (arena.dry) return DDPM3D_OK;
    prof_begin(3, 0.0);
    const int r = gn_finalize_only(g, s);
    prof_end();
    return r;
  }

  int gn(GnArgs& g) {
    g.B = B;
    g.Z = Z;
    // 1: finalize -> apply; 2: also conv -> finalize and apply -> conv (the whole conv / GroupNorm chain; not under the profiler's brackets)
    g.pdl = (ctx->pdl >= 2 && ctx->profile) ? 1 : ctx->pdl;
    g.stream_allowed = ctx->gn_stream;
    g.stream_min_mb = ctx->gn_stream_mb;
    const int Ctot = g.C[0] + g.C[1];
    g.n_chunks = gn_chunks((int64_t)Z * g.H * g.W);
    g.partials = (double*)arena.alloc((size_t)B * g.n_chunks * 64 * sizeof(double));
    g.ab = (float*)arena.alloc((size_t)B * 2 * Ctot * sizeof(float));
    double* sums = nullptr;
    double* gathered = nullptr;
    if (zp) {
      sums = (double*)arena.alloc((size_t)B * 64 * sizeof(double));
      gathered = (double*)arena.alloc((size_t)ctx->slab.world * B * 64 * sizeof(double));
      launches += 2;
      if (arena.dry && g.out_zpad && p2p()) launches += 2;  // (the live path counts them in halo_begin / halo_end)
    }
    const bool have_cs = !g.pre_add && g.chsum[0] && (g.C[1] == 0 || g.chsum[1]);
    const bool fused = !zp && have_cs;
    launches += fused ? 2 : 3;  // (statistics,) finalize, apply
    if (arena.dry) return DDPM3D_OK;
    const double n = (double)B * Z * g.H * g.W * Ctot;
    const double in_b = is_half_dt(g.dt) ? 2 : 4, out_b = (is_half_dt(g.dt) && !g.out_f32) ? 2 : 4;
    const double scale = g.resample == RS_POOL ? 0.25 : (g.resample == RS_UP ? 4.0 : 1.0);
    prof_begin(4, n * ((have_cs ? 1 : 2) * in_b + out_b * scale));
    int r;
    if (fused) {
      r = gn_forward_chsum(g, s);
    } else if (zp) {  // z-slab sharding: statistics span all ranks (fp64 sums exchanged, summed in rank order)
      r = have_cs ? gn_chsum_local(g, sums, s) : gn_stats_local(g, sums, s);
      const bool peer = p2p() && B * 64 <= SLAB_GATHER_DOUBLES;
      if (r == DDPM3D_OK) {
        prof_end();
        prof_begin(11, (double)ctx->slab.world * B * 64 * sizeof(double));
        if (peer) {  // every rank stores its sums into every mailbox; the finalize kernel waits for the flags
          r = comm_stats_push(ctx->slab, sums, B * 64, s);
          g.gathered = comm_stats_slots(ctx->slab);
          g.gather_stride = SLAB_GATHER_DOUBLES;
          g.gather_parity_stride = (int64_t)SLAB_MAX_RANKS * SLAB_GATHER_DOUBLES;
          g.gather_flags = comm_stats_flags(ctx->slab);
          g.gather_seq = comm_stats_seq(ctx->slab);
        } else {
          r = comm_allgather_f64(ctx->slab, sums, gathered, (size_t)B * 64, s);
          g.gathered = gathered;
        }
        prof_end();
      }
      void *peer_lo = nullptr, *peer_hi = nullptr;
      const bool peer_halo = r == DDPM3D_OK && g.out_zpad && p2p();
      if (peer_halo) {
        const int Ho = g.resample == RS_POOL ? g.H / 2 : (g.resample == RS_UP ? 2 * g.H : g.H);
        const int Wo = g.resample == RS_POOL ? g.W / 2 : (g.resample == RS_UP ? 2 * g.W : g.W);
        r = halo_begin(g.out, (size_t)Ho * Wo * Ctot * (size_t)out_b, &peer_lo, &peer_hi);
        g.peer_halo[0] = peer_lo;
        g.peer_halo[1] = peer_hi;
      }
      if (r == DDPM3D_OK) {
        prof_begin(4, 0.0);
        g.world = ctx->slab.world;
        g.inv_count_global = 1.0 / ((double)ctx->slab.z_total * g.H * g.W * (Ctot / 32));
        r = gn_finalize_apply(g, s);
        prof_end();
        if (peer_halo && r == DDPM3D_OK) r = halo_end(2.0 * B * g.H * g.W * Ctot * out_b * scale);
        return r;
      }
    } else {
      int dummy = 0;
      r = gn_forward(g, s, &dummy);
    }
    prof_end();
    return r;
  }
};

int run_res(Run& R, const Layer& L, const Act* src, int nsrc, Act* out) {
  ddpm3d_ctx* ctx = R.ctx;
  const int dt = ctx->dt, dts = ctx->dts;
  int Cin = 0;
  for (int i = 0; i < nsrc; ++i) Cin += src[i].C;
  DD_CHECK(Cin == L.cin, DDPM3D_ERR_STATE, "internal: ResBlock input channel mismatch at " + L.prefix);
  const int H = src[0].H, W = src[0].W;
  const int Ho = L.down ? H / 2 : (L.up ? 2 * H : H), Wo = L.down ? W / 2 : (L.up ? 2 * W : W);
  DD_CHECK(!L.down || (H % 2 == 0 && W % 2 == 0), DDPM3D_ERR_ARG, "H and W must be divisible by 2^(levels-1)");
  DD_CHECK(!(L.skip_conv && (L.up || L.down)), DDPM3D_ERR_ARG, "internal: resampling ResBlock with a skip conv");
  DD_CHECK(L.skip_conv || nsrc == 1, DDPM3D_ERR_ARG, "identity skip over a channel concat is not supported");
  out->C = L.cout; out->H = Ho; out->W = Wo;
  out->p = R.arena.alloc(R.act_bytes(Ho, Wo, L.cout));
  float* out_cs = R.alloc_chsum(L.cout);
  const size_t mark = R.arena.off;

  // in_layers: GN32 -> SiLU (-> h_upd)                                   unet.py:237-244
  void* h1 = R.arena.alloc(R.conv_in_bytes(Ho, Wo, Cin, ctx->esz));
  GnArgs g{};
  g.out_zpad = R.zp;
  g.dt = dts;
  g.dt_out = dt;
  for (int i = 0; i < nsrc; ++i) {
    g.src[i] = src[i].p; g.C[i] = src[i].C;
    g.chsum[i] = src[i].chsum; g.chsum_bias[i] = src[i].cs_bias; g.chsum_P[i] = src[i].cs_slots;
  }
  g.H = H; g.W = W;
  g.gamma = L.gn1_g; g.beta = L.gn1_b;
  g.silu = 1;
  g.resample = L.down ? RS_POOL : (L.up ? RS_UP : RS_NONE);
  g.out = h1;
  DD_TRY(R.gn(g));
  DD_TRY(R.halo(h1, Ho, Wo, Cin, ctx->esz));
  // in_layers[-1]: conv 3x3x3
  void* h2 = R.arena.alloc(R.act_bytes(Ho, Wo, L.cout));
  ConvArgs c{};
  c.dt = dt;
  c.dt_io = dts;
  c.main = {h1, Cin};
  c.taps = L.c1.taps;
  c.w = L.c1.w; c.bias = L.c1.bias;
  c.out = h2; c.Ho = Ho; c.Wo = Wo; c.Cout = L.cout;
  c.chsum_out = R.alloc_chsum(L.cout);
  DD_TRY(R.conv(c));
  // out_layers: GN32 (FiLM | +emb) -> SiLU -> conv                       unet.py:245-255
  void* h3 = R.arena.alloc(R.conv_in_bytes(Ho, Wo, L.cout, ctx->esz));
  GnArgs g2{};
  g2.out_zpad = R.zp;
  g2.dt = dts;
  g2.dt_out = dt;
  g2.src[0] = h2; g2.C[0] = L.cout;
  g2.chsum[0] = c.chsum_written ? c.chsum_out : nullptr;
  g2.chsum_bias[0] = L.c1.bias;
  g2.H = Ho; g2.W = Wo;
  g2.gamma = L.gn2_g; g2.beta = L.gn2_b;
  if (ctx->cfg.use_scale_shift_norm) { g2.film = R.emb_out ? R.emb_out + L.emb_off : nullptr; g2.film_stride = ctx->emb_rows_total; }
  else { g2.pre_add = R.emb_out ? R.emb_out + L.emb_off : nullptr; g2.pre_stride = ctx->emb_rows_total; }
  if (R.arena.dry) { g2.film = nullptr; g2.pre_add = nullptr; }
  g2.silu = 1;
  g2.out = h3;
  DD_TRY(R.gn(g2));
  DD_TRY(R.halo(h3, Ho, Wo, L.cout, ctx->esz));
  ConvArgs c2{};
  c2.dt = dt;
  c2.dt_io = dts;
  c2.main = {h3, L.cout};
  c2.taps = L.c2.taps;
  c2.w = L.c2.w; c2.bias = L.c2.bias;
  c2.out = out->p; c2.Ho = Ho; c2.Wo = Wo; c2.Cout = L.cout;
  if (L.skip_conv) {  // skip_connection (1x1x1) folded into the same accumulation, reading the concat halves in place
    c2.n_extra = nsrc;
    for (int i = 0; i < nsrc; ++i) c2.extra[i] = {src[i].p, src[i].C};
  } else if (!L.up && !L.down && ctx->fold_identity && is_half_dt(dt)) {
    c2.n_extra = 1;  // Identity skip as a unit-weight 1x1x1 source: no residual traffic in the epilogue
    c2.extra[0] = {src[0].p, src[0].C};
    c2.extra_is_identity = 1;
  } else {
    c2.residual = src[0].p;
    c2.res_mode = L.down ? RES_POOL : (L.up ? RES_UP : RES_SAME);
    if (!L.up && !L.down) c2.w_ld = L.c2.taps * L.cout + L.cout;  // skip the appended unit block
  }
  c2.chsum_out = out_cs;
  DD_TRY(R.conv(c2));
  out->chsum = c2.chsum_written ? out_cs : nullptr;
  out->cs_bias = L.c2.bias;
  R.arena.off = mark;
  return DDPM3D_OK;
}

int run_attn(Run& R, const Layer& L, const Act& x, Act* out) {
  ddpm3d_ctx* ctx = R.ctx;
  const int dt = ctx->dt, dts = ctx->dts, C = L.cin, H = x.H, W = x.W;
  DD_CHECK(x.C == C, DDPM3D_ERR_STATE, "internal: attention channel mismatch");
  out->C = C; out->H = H; out->W = W;
  out->p = R.arena.alloc(R.act_bytes(H, W, C));
  const size_t mark = R.arena.off;
  void* n = R.arena.alloc(R.act_bytes(H, W, C));
  GnArgs g{};
  g.dt = dts; g.dt_out = dt; g.src[0] = x.p; g.C[0] = C; g.H = H; g.W = W; g.gamma = L.gn1_g; g.beta = L.gn1_b; g.silu = 0; g.out = n;
  g.chsum[0] = x.chsum; g.chsum_bias[0] = x.cs_bias; g.chsum_P[0] = x.cs_slots;
  DD_TRY(R.gn(g));
  void* qkv = R.arena.alloc(R.act_bytes(H, W, 3 * C));
  ConvArgs c{};
  c.dt = dt; c.main = {n, C}; c.taps = 1; c.w = L.c1.w; c.bias = L.c1.bias; c.out = qkv; c.Ho = H; c.Wo = W; c.Cout = 3 * C;
  DD_TRY(R.conv(c));
  // z-slab sharding: the queries stay local, keys and values of the whole volume are all-gathered
  // along T.  The packed qkv rows travel as they are (the unused third, the other ranks' queries, is cheaper than a
  // repack at these resolutions); rank order = z order, so the gathered tensor is the global [T][3C] token matrix.
  const int world = R.zp ? ctx->slab.world : 1;
  const int T_local = R.Z * H * W, T_all = T_local * world;
  const void* qkv_all = qkv;
  int q_begin = 0;
  if (R.zp) {
    DD_CHECK(ctx->slab.z_total == world * R.Z && ctx->slab.z_begin == ctx->slab.rank * R.Z, DDPM3D_ERR_ARG,
             "z-slab sharding with attention blocks needs equal slabs in rank order (z_total = world * Z)");
    void* gathered = R.arena.alloc((size_t)world * R.act_bytes(H, W, 3 * C));
    ++R.launches;
    if (!R.arena.dry) {
      const size_t per_b = (size_t)T_local * 3 * C * ctx->esz;
      R.prof_begin(10, (double)R.B * per_b * world);
      const int r = comm_allgather_slabs(ctx->slab, qkv, gathered, R.B, per_b, per_b, per_b * world, R.s);
      R.prof_end();
      DD_TRY(r);
    }
    qkv_all = gathered;
    q_begin = ctx->slab.rank * T_local;
  }
  void* a = R.arena.alloc(R.act_bytes(H, W, C));
  const size_t at_bytes = ctx->conv_path == 1 ? 0 : attention_tc_scratch_bytes(dt, R.B, T_all, C, L.heads);
  void* at_scratch = at_bytes ? R.arena.alloc(at_bytes) : nullptr;
  R.launches += at_bytes ? 2 : 1;
  if (!R.arena.dry) {
    R.prof_begin(7, 4.0 * R.B * (double)T_local * T_all * C);
    const int r = attention_k(dt, qkv_all, a, R.B, T_all, C, L.heads, ctx->cfg.use_new_attention_order, at_scratch, at_bytes,
                              R.s, q_begin, T_local);
    R.prof_end();
    DD_TRY(r);
  }
  ConvArgs p{};
  p.dt = dt; p.dt_io = dts; p.main = {a, C}; p.taps = 1; p.w = L.c2.w; p.bias = L.c2.bias; p.out = out->p; p.Ho = H; p.Wo = W; p.Cout = C;
  p.residual = x.p; p.res_mode = RES_SAME;
  DD_TRY(R.conv(p));
  R.arena.off = mark;
  return DDPM3D_OK;
}

int run_conv_layer(Run& R, const Layer& L, const Act& x, Act* out) {
  ddpm3d_ctx* ctx = R.ctx;
  const int dt = ctx->dts;  // reads a block input / the packed network input: runs in the storage format
  DD_CHECK(x.C == L.cin, DDPM3D_ERR_STATE, "internal: conv channel mismatch at " + L.prefix);
  DD_CHECK(!R.zp || L.prefix == "input_blocks.0.0", DDPM3D_ERR_ARG,
           "z-slab sharding needs resblock_updown=True (bare Downsample / Upsample convs read un-haloed tensors)");
  int Ho = x.H, Wo = x.W;
  const void* in = x.p;
  size_t mark = 0;
  if (L.kind == L_UPCONV) {  // Upsample(use_conv=True): nearest x2 then conv (unet.py:100-108)
    Ho = 2 * x.H; Wo = 2 * x.W;
  } else if (L.stride_hw == 2) {  // Downsample(use_conv=True): stride (1,2,2) (unet.py:129-133)
    DD_CHECK(x.H % 2 == 0 && x.W % 2 == 0, DDPM3D_ERR_ARG, "H and W must be divisible by 2^(levels-1)");
    Ho = x.H / 2; Wo = x.W / 2;
  }
  out->C = L.cout; out->H = Ho; out->W = Wo;
  out->p = R.arena.alloc(R.act_bytes(Ho, Wo, L.cout));
  // channel sums of the output for the GroupNorm that reads it: the tensor-core stem has one slot per CTA
  const size_t cs_floats = (size_t)R.B * 6 * chsum_slots() * L.cout * 2;
  float* out_cs = (float*)R.arena.alloc(cs_floats * sizeof(float));
  mark = R.arena.off;
  if (L.kind == L_UPCONV) {
    void* u = R.arena.alloc(R.act_bytes(Ho, Wo, L.cin));
    ++R.launches;
    if (!R.arena.dry) {
      R.prof_begin(8, (double)R.act_bytes(x.H, x.W, L.cin) * 5.0);
      const int r = resample_hw(dt, x.p, u, R.B, R.Z, x.H, x.W, L.cin, RS_UP, R.s);
      R.prof_end();
      DD_TRY(r);
    }
    in = u;
  }
  ConvArgs c{};
  c.dt = dt; c.main = {in, L.cin}; c.taps = L.c1.taps; c.stride_hw = L.kind == L_UPCONV ? 1 : L.stride_hw;
  c.w = L.c1.w; c.bias = L.c1.bias; c.out = out->p; c.Ho = Ho; c.Wo = Wo; c.Cout = L.cout;
  c.chsum_out = out_cs;
  DD_TRY(R.conv(c));
  if (c.chsum_written) {
    out->chsum = out_cs;
    out->cs_bias = L.c1.bias;
    out->cs_slots = conv_stem_eligible(c) ? conv_stem_chsum_slots(c) : 0;  // (0 = one per SM: the tcgen05 conv kernels)
  }
  R.arena.off = mark;
  return DDPM3D_OK;
}

int run_block(Run& R, const std::vector<Layer>& blk, const Act* src, int nsrc, Act* out) {
  Act cur[2] = {src[0], nsrc > 1 ? src[1] : Act{}};
  int n = nsrc;
  for (const Layer& L : blk) {
    Act o;
    if (L.kind == L_RES) DD_TRY(run_res(R, L, cur, n, &o));
    else {
      DD_CHECK(n == 1, DDPM3D_ERR_STATE, "internal: concat input to a non-ResBlock layer");
      if (L.kind == L_ATTN) DD_TRY(run_attn(R, L, cur[0], &o));
      else DD_TRY(run_conv_layer(R, L, cur[0], &o));
    }
    cur[0] = o;
    n = 1;
  }
  *out = cur[0];
  return DDPM3D_OK;
}

// SuperResModel_noatt.forward (unet.py:1687-1694) + UNetModel_noatt.forward (unet.py:1015-1044)
int forward_impl(ddpm3d_ctx* ctx, Run& R, const float* x, const float* low, const float* t, const int64_t* y, float* out,
                 int H, int W) {
  const int B = R.B, Z = R.Z;
  // profiling: an empty bracket, so the reader knows what two back-to-back event records cost in this launch mode
  R.prof_begin(12, 0.0);
  R.prof_end();
  // time_embed + every emb_layers Linear (unet.py:1029-1033, 199-205)
  float* emb_silu = (float*)R.arena.alloc((size_t)B * ctx->ted * sizeof(float));
  R.emb_out = (float*)R.arena.alloc((size_t)B * std::max(ctx->emb_rows_total, 1) * sizeof(float));
  R.launches += 2;
  if (!R.arena.dry) {
    EmbArgs e{};
    e.t = t; e.y = y; e.B = B; e.model_channels = ctx->cfg.model_channels; e.ted = ctx->ted;
    e.w0 = ctx->te_w0; e.b0 = ctx->te_b0; e.w2 = ctx->te_w2; e.b2 = ctx->te_b2;
    e.label_emb = ctx->cfg.num_classes > 0 ? ctx->label_emb : nullptr;
    DD_CHECK(!e.label_emb || y, DDPM3D_ERR_ARG, "class-conditional model needs y (unet.py:1024-1026)");
    e.freqs = ctx->d_freqs;
    e.emb_silu = emb_silu;
    e.w_all = ctx->emb_w_all; e.b_all = ctx->emb_b_all; e.rows_total = ctx->emb_rows_total;
    e.emb_out = R.emb_out;
    int dummy = 0;
    R.prof_begin(5, 0);
    const int r = embedding_forward(e, R.s, &dummy);
    R.prof_end();
    DD_TRY(r);
  }
  // cat([x, low_res], 1).type(dtype)
  Act h;
  const int Cx = ctx->cfg.in_channels, Cstem = ctx->stem_cin();
  h.C = Cstem; h.H = H; h.W = W;
  h.p = R.arena.alloc(R.conv_in_bytes(H, W, Cstem, ctx->esz));
  ++R.launches;
  if (!R.arena.dry) {
    DD_CHECK((low != nullptr) == (ctx->cfg.unconditional == 0), DDPM3D_ERR_ARG,
             "low_res must be given to a conditional model and only to it (unet.py:1687-1694)");
    void *peer_lo = nullptr, *peer_hi = nullptr;
    const bool peer = R.p2p() && Cx == 1 && low;
    if (peer) DD_TRY(R.halo_begin(h.p, (size_t)H * W * Cstem * ctx->esz, &peer_lo, &peer_hi));
    R.prof_begin(8, (double)B * Z * H * W * Cstem * (4.0 + ctx->esz));
    const int r = Cx == 1 && low ? pack_input(ctx->dts, x, low, h.p, B, Z, (int64_t)H * W, R.zp, R.s, peer_lo, peer_hi)
                                 : pack_input_planar(ctx->dts, x, low, Cx, h.p, B, Z, (int64_t)H * W, R.zp, R.s);
    R.prof_end();
    DD_TRY(r);
    if (peer) DD_TRY(R.halo_end((double)B * 2 * H * W * Cstem * ctx->esz));
    else if (R.p2p()) DD_TRY(R.halo_nccl(h.p, H, W, Cstem, ctx->esz));
  }
  if (!R.p2p()) DD_TRY(R.halo(h.p, H, W, Cstem, ctx->esz));
  std::vector<Act> skips;
  for (auto& blk : ctx->input_blocks) {
    Act o;
    DD_TRY(run_block(R, blk, &h, 1, &o));
    h = o;
    skips.push_back(h);
  }
  {
    Act o;
    DD_TRY(run_block(R, ctx->middle, &h, 1, &o));
    h = o;
  }
  for (auto& blk : ctx->output_blocks) {
    Act src[2] = {h, skips.back()};
    skips.pop_back();
    DD_CHECK(src[0].H == src[1].H && src[0].W == src[1].W, DDPM3D_ERR_STATE, "internal: skip geometry mismatch");
    Act o;
    DD_TRY(run_block(R, blk, src, 2, &o));
    h = o;
  }
  DD_CHECK(h.H == H && h.W == W && h.C == ctx->out_norm_ch, DDPM3D_ERR_STATE, "internal: output geometry mismatch");
  // h.type(x.dtype); out = GN32 -> SiLU -> conv (fp32)                 unet.py:1043-1044
  if (ctx->head_tc && is_half_dt(ctx->dts) && !R.zp && ctx->conv_path != 1 && !ctx->two_d() &&
      conv_head_tc_eligible(ctx->dts, h.C, ctx->cfg.out_channels)) {
    // 16-bit modes: one kernel applies the GroupNorm affine + SiLU while staging and runs the conv on tcgen05
    GnArgs g{};
    g.chsum[0] = h.chsum;
    g.chsum_bias[0] = h.cs_bias;
    g.chsum_P[0] = h.cs_slots;
    g.dt = ctx->dts; g.src[0] = h.p; g.C[0] = h.C; g.H = H; g.W = W; g.gamma = ctx->out_gn_g; g.beta = ctx->out_gn_b; g.silu = 1;
    DD_TRY(R.gn_stats(g));
    ++R.launches;
    if (!R.arena.dry) {
      R.prof_begin(9, 2.0 * B * Z * H * W * (double)ctx->cfg.out_channels * 27.0 * h.C);
      const int r = conv_head_tc(ctx->dts, h.p, g.ab, (const float*)ctx->out_conv.w, ctx->out_conv.bias, out, B, Z, H, W, h.C,
                                 ctx->cfg.out_channels, R.s);
      R.prof_end();
      DD_TRY(r);
    }
    return DDPM3D_OK;
  }
  float* hn = (float*)R.arena.alloc(R.conv_in_bytes(H, W, h.C, sizeof(float)));
  GnArgs g{};
  g.out_zpad = R.zp;
  g.chsum[0] = h.chsum;
  g.chsum_bias[0] = h.cs_bias;
  g.chsum_P[0] = h.cs_slots;
  g.dt = ctx->dts; g.src[0] = h.p; g.C[0] = h.C; g.H = H; g.W = W; g.gamma = ctx->out_gn_g; g.beta = ctx->out_gn_b; g.silu = 1;
  g.out = hn; g.out_f32 = 1;
  DD_TRY(R.gn(g));
  DD_TRY(R.halo(hn, H, W, h.C, sizeof(float)));
  ConvArgs c{};
  c.dt = DDPM3D_FP32; c.main = {hn, h.C}; c.taps = ctx->out_conv.taps; c.w = ctx->out_conv.w; c.bias = ctx->out_conv.bias;
  c.out = out; c.out_planar_f32 = 1; c.Ho = H; c.Wo = W; c.Cout = ctx->cfg.out_channels;
  DD_TRY(R.conv(c));
  return DDPM3D_OK;
}

int check_geometry(ddpm3d_ctx* ctx, int B, int Z, int H, int W) {
  DD_CHECK(ctx->finalized, DDPM3D_ERR_STATE, "weights not finalized (call ddpm3d_finalize_weights first)");
  DD_CHECK(B >= 1 && Z >= 1 && H >= 1 && W >= 1, DDPM3D_ERR_ARG, "bad geometry");
  DD_CHECK(!ctx->two_d() || Z == 1, DDPM3D_ERR_ARG, "a dims = 2 model takes (B, C, H, W) images: Z must be 1");
  const int f = 1 << (ctx->cfg.n_levels - 1);
  DD_CHECK(H % f == 0 && W % f == 0, DDPM3D_ERR_ARG, "H and W must be divisible by 2^(levels-1)");
  DD_CHECK(((int64_t)Z * H * W) % 4 == 0, DDPM3D_ERR_ARG, "Z*H*W must be a multiple of 4");
  // a sharded call takes a proper part of the volume: a stale set_slab (or an un-sharded patch while the sharded path
  // is still switched on) is refused instead of silently exchanging halos with the other ranks
  if (ctx->slab.active())
    DD_CHECK(ctx->slab.z_total > Z && ctx->slab.z_begin + Z <= ctx->slab.z_total, DDPM3D_ERR_STATE,
             "z-slab sharding is on but this call is not a slab of the volume set by ddpm3d_set_slab(z_begin, z_total): "
             "set the slab for this volume, or switch sharding off with ddpm3d_set_slab(ctx, 0, 0)");
  return DDPM3D_OK;
}

int64_t dry_bytes(ddpm3d_ctx* ctx, int B, int Z, int H, int W, int* launches) {
  Run R{ctx, B, Z, nullptr};
  R.zp = ctx->slab.active() ? 1 : 0;
  R.arena.dry = true;
  if (forward_impl(ctx, R, nullptr, nullptr, nullptr, nullptr, nullptr, H, W) != DDPM3D_OK) return -1;
  if (launches) *launches = R.launches;
  return (int64_t)R.arena.peak;
}

int ensure_workspace(ddpm3d_ctx* ctx, int B, int Z, int H, int W) {
  ctx->splitk_need = 0;
  const int64_t need = dry_bytes(ctx, B, Z, H, W, nullptr);
  if (need < 0) return DDPM3D_ERR_ARG;
  if (ctx->splitk_need > ctx->splitk_cap) {
    for (auto& kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
    ctx->graphs.clear();
    ctx->graph_launches.clear();
    DD_CUDA(cudaDeviceSynchronize());
    if (ctx->splitk_buf) DD_CUDA(cudaFree(ctx->splitk_buf));
    ctx->splitk_buf = nullptr;
    ctx->splitk_cap = 0;
    DD_CUDA(cudaMalloc((void**)&ctx->splitk_buf, ctx->splitk_need));
    ctx->splitk_cap = ctx->splitk_need;
  }
  if ((size_t)need > ctx->ws_cap) {
    // graphs captured on the old workspace are stale
    for (auto& kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
    ctx->graphs.clear();
    ctx->graph_launches.clear();
    DD_CUDA(cudaDeviceSynchronize());
    if (ctx->ws) DD_CUDA(cudaFree(ctx->ws));
    ctx->ws = nullptr;
    ctx->ws_cap = 0;
    DD_CUDA(cudaMalloc((void**)&ctx->ws, (size_t)need));
    ctx->ws_cap = (size_t)need;
  }
  return DDPM3D_OK;
}

// z-slab sharding over peer-mapped memory: (re)map the neighbours' workspaces.  Collective over the slab ranks: every
// rank reaches it at the start of every sharded call (same call sequence on all ranks).
int sync_peers(ddpm3d_ctx* ctx, cudaStream_t s) {
  if (!ctx->slab.active() || !ctx->slab.p2p || !ctx->slab_p2p) return DDPM3D_OK;
  bool changed = false;
  DD_TRY(comm_peer_sync_ws(&ctx->slab, ctx->ws, s, &changed));
  if (changed) {  // captured graphs hold the old peer addresses
    for (auto& kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
    ctx->graphs.clear();
    ctx->graph_launches.clear();
  }
  return DDPM3D_OK;
}

// A sharded call may be captured in a CUDA graph when nothing in it goes through NCCL: peer path, equal slabs, no
// attention blocks (their K/V all-gather is an ncclAllGather)
bool slab_capturable(const ddpm3d_ctx* ctx, int B, int Z) {
  const SlabComm& c = ctx->slab;
  if (!ctx->slab_p2p || !c.halo_p2p() || c.z_total != c.world * Z || c.z_begin != c.rank * Z || B * 64 > SLAB_GATHER_DOUBLES) return false;
  if (ctx->cfg.in_channels != 1 || ctx->cfg.unconditional) return false;  // (the planar pack kernel exchanges through NCCL)
  auto has_attn = [](const std::vector<Layer>& blk) {
    for (const Layer& L : blk)
      if (L.kind == L_ATTN) return true;
    return false;
  };
  for (auto& blk : ctx->input_blocks)
    if (has_attn(blk)) return false;
  if (has_attn(ctx->middle)) return false;
  for (auto& blk : ctx->output_blocks)
    if (has_attn(blk)) return false;
  return true;
}

int forward_launch(ddpm3d_ctx* ctx, const float* x, const float* low, const float* t, const int64_t* y, float* out, int B,
                   int Z, int H, int W, cudaStream_t s, int* launches) {
  Run R{ctx, B, Z, s};
  R.zp = ctx->slab.active() ? 1 : 0;
  R.arena.dry = false;
  R.arena.base = ctx->ws;
  R.arena.cap = ctx->ws_cap;
  DD_TRY(forward_impl(ctx, R, x, low, t, y, out, H, W));
  DD_CHECK(!R.arena.overflow, DDPM3D_ERR_STATE, "internal: workspace overflow");
  if (launches) *launches += R.launches;
  return DDPM3D_OK;
}

// Runs `body` (which enqueues work on `s`) either directly or through a cached CUDA graph.
template <typename F>
int run_graphed(ddpm3d_ctx* ctx, const GraphKey& key0, cudaStream_t s, F&& body) {
  GraphKey key = key0;  // a sharded step bakes the slab position and the peer addresses in
  if (ctx->slab.active()) key.i[2] = ((int64_t)1 << 62) | ((int64_t)ctx->slab.z_begin << 31) | (int64_t)ctx->slab.z_total;
  // NCCL calls are issued eagerly, in rank-identical order; the peer path has none and replays like any other step
  if (!ctx->use_graph || ctx->profile == 1 || (ctx->slab.active() && !slab_capturable(ctx, key.B, key.Z))) {
    int n = 0;
    DD_TRY(body(s, &n));
    ctx->launches += n;
    return DDPM3D_OK;
  }
  // the caller's stream may be the legacy default stream, which cannot be captured: capture on an
  // internal stream, replay on the caller's
  if (!ctx->cap_stream) DD_CUDA(cudaStreamCreateWithFlags(&ctx->cap_stream, cudaStreamNonBlocking));
  auto it = ctx->graphs.find(key);
  if (it == ctx->graphs.end()) {
    if (ctx->graphs.size() >= 16) {
      for (auto& kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
      ctx->graphs.clear();
      ctx->graph_launches.clear();
    }
    cudaGraph_t graph = nullptr;
    int n = 0;
    DD_CUDA(cudaStreamBeginCapture(ctx->cap_stream, cudaStreamCaptureModeRelaxed));
    const int r = body(ctx->cap_stream, &n);
    const cudaError_t e = cudaStreamEndCapture(ctx->cap_stream, &graph);
    if (r != DDPM3D_OK) {
      if (graph) cudaGraphDestroy(graph);
      return r;
    }
    if (e != cudaSuccess) {
      set_error(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
      return DDPM3D_ERR_CUDA;
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t e2 = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e2 != cudaSuccess) {
      set_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e2));
      return DDPM3D_ERR_CUDA;
    }
    it = ctx->graphs.emplace(key, exec).first;
    ctx->graph_launches[key] = n;
  }
  DD_CUDA(cudaGraphLaunch(it->second, s));
  ctx->launches += ctx->graph_launches[key];
  return DDPM3D_OK;
}

int ensure_loop_buffers(ddpm3d_ctx* ctx, int B, int64_t n_vox) {
  DD_CHECK(ctx->d_table != nullptr, DDPM3D_ERR_STATE, "schedule not set (call ddpm3d_set_schedule first)");
  if (B > ctx->loop_B) {
    DD_CUDA(cudaDeviceSynchronize());
    if (ctx->d_tmodel) cudaFree(ctx->d_tmodel);
    if (ctx->d_tindex) cudaFree(ctx->d_tindex);
    DD_CUDA(cudaMalloc((void**)&ctx->d_tmodel, (size_t)B * sizeof(float)));
    DD_CUDA(cudaMalloc((void**)&ctx->d_tindex, (size_t)B * sizeof(int32_t)));
    ctx->loop_B = B;
  }
  if (!ctx->d_counter) DD_CUDA(cudaMalloc((void**)&ctx->d_counter, 4 * sizeof(int32_t)));
  const size_t mo = (size_t)B * n_vox * 2 * sizeof(float);
  if (mo > ctx->mo_cap) {
    DD_CUDA(cudaDeviceSynchronize());
    if (ctx->d_mo) cudaFree(ctx->d_mo);
    DD_CUDA(cudaMalloc((void**)&ctx->d_mo, mo));
    ctx->mo_cap = mo;
  }
  const size_t im = (size_t)B * n_vox * sizeof(float);
  if (im > ctx->img_cap) {
    DD_CUDA(cudaDeviceSynchronize());
    if (ctx->d_img) cudaFree(ctx->d_img);
    DD_CUDA(cudaMalloc((void**)&ctx->d_img, im));
    ctx->img_cap = im;
  }
  return DDPM3D_OK;
}

bool learned_var(int v) { return v == DDPM3D_VAR_LEARNED || v == DDPM3D_VAR_LEARNED_RANGE; }

int check_sampler_shapes(ddpm3d_ctx* ctx) {
  DD_CHECK(ctx->cfg.in_channels == 1 && !ctx->cfg.unconditional, DDPM3D_ERR_ARG,
           "the fused UNet + update entry points serve the one-channel conditional model; other model classes go "
           "through ddpm3d_unet_forward + ddpm3d_p_sample_update");
  const int need = learned_var(ctx->var_type) ? 2 : 1;
  DD_CHECK(ctx->cfg.out_channels == need, DDPM3D_ERR_ARG,
           "model out_channels does not match the variance type (gaussian_diffusion.py:262-263)");
  return DDPM3D_OK;
}

// scratch for the single-kernel test entry points
struct Scratch {
  void* p = nullptr;
  size_t cap = 0;
  int get(size_t bytes, void** out) {
    if (bytes > cap) {
      DD_CUDA(cudaDeviceSynchronize());
      if (p) cudaFree(p);
      p = nullptr;
      cap = 0;
      DD_CUDA(cudaMalloc(&p, bytes));
      cap = bytes;
    }
    *out = p;
    return DDPM3D_OK;
  }
};
Scratch g_scratch;

}  // namespace
}  // namespace ddpm3d

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* ddpm3d_last_error(void) { return g_error.c_str(); }
int ddpm3d_abi_version(void) { return DDPM3D_ABI_VERSION; }

int ddpm3d_create(const ddpm3d_config* cfg, ddpm3d_ctx** out) {
  DD_CHECK(cfg && out, DDPM3D_ERR_ARG, "ddpm3d_create: null argument");
  DD_CHECK(cfg->precision >= DDPM3D_FP32 && cfg->precision <= DDPM3D_BF16_STRICT, DDPM3D_ERR_ARG, "config: bad precision");
  std::unique_ptr<ddpm3d_ctx> ctx(new ddpm3d_ctx());
  ctx->cfg = *cfg;
  ctx->dt = cfg->precision == DDPM3D_BF16_STRICT ? DDPM3D_BF16 : cfg->precision;
  ctx->dts = cfg->precision == DDPM3D_BF16 ? DDPM3D_FP16 : ctx->dt;
  ctx->esz = is_half_dt(ctx->dt) ? 2 : 4;
  DD_TRY(build_topology(ctx.get()));
  *out = ctx.release();
  return DDPM3D_OK;
}

void ddpm3d_destroy(ddpm3d_ctx* ctx) {
  if (!ctx) return;
  if (ctx->device >= 0) {
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto& kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
    for (void* p : ctx->dev_allocs) cudaFree(p);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->d_table) cudaFree(ctx->d_table);
    if (ctx->d_tmodel) cudaFree(ctx->d_tmodel);
    if (ctx->d_tindex) cudaFree(ctx->d_tindex);
    if (ctx->d_counter) cudaFree(ctx->d_counter);
    if (ctx->d_mo) cudaFree(ctx->d_mo);
    if (ctx->d_img) cudaFree(ctx->d_img);
    if (ctx->d_freqs) cudaFree(ctx->d_freqs);
    if (ctx->splitk_buf) cudaFree(ctx->splitk_buf);
    comm_destroy(&ctx->slab);
    if (ctx->cap_stream) cudaStreamDestroy(ctx->cap_stream);
    for (auto& e : ctx->prof) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  }
  delete ctx;
}

int ddpm3d_param_count(const ddpm3d_ctx* ctx) { return ctx ? (int)ctx->params.size() : 0; }

int ddpm3d_param_info(const ddpm3d_ctx* ctx, int i, const char** key, int64_t shape[8], int* ndim) {
  DD_CHECK(ctx && i >= 0 && i < (int)ctx->params.size(), DDPM3D_ERR_ARG, "param_info: index out of range");
  const Param& p = ctx->params[i];
  if (key) *key = p.key.c_str();
  if (ndim) *ndim = (int)p.shape.size();
  if (shape)
    for (size_t d = 0; d < p.shape.size() && d < 8; ++d) shape[d] = p.shape[d];
  return DDPM3D_OK;
}

int ddpm3d_load_tensor(ddpm3d_ctx* ctx, const char* key, const float* data, const int64_t* shape, int ndim) {
  DD_CHECK(ctx && key && data && shape, DDPM3D_ERR_ARG, "load_tensor: null argument");
  DD_CHECK(!ctx->finalized, DDPM3D_ERR_STATE, "load_tensor: weights already finalized");
  auto it = ctx->index.find(key);
  DD_CHECK(it != ctx->index.end(), DDPM3D_ERR_MISSING, std::string("unexpected key in state_dict: ") + key);
  Param& p = ctx->params[it->second];
  bool same = (int)p.shape.size() == ndim;
  for (int d = 0; same && d < ndim; ++d) same = p.shape[d] == shape[d];
  DD_CHECK(same, DDPM3D_ERR_ARG, std::string("size mismatch for ") + key);
  cudaPointerAttributes attr{};
  const bool on_device = cudaPointerGetAttributes(&attr, data) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
  cudaGetLastError();
  p.host.resize((size_t)p.numel());
  if (on_device) DD_CUDA(cudaMemcpy(p.host.data(), data, p.host.size() * sizeof(float), cudaMemcpyDeviceToHost));
  else std::memcpy(p.host.data(), data, p.host.size() * sizeof(float));
  p.loaded = true;
  return DDPM3D_OK;
}

int ddpm3d_finalize_weights(ddpm3d_ctx* ctx, int device) {
  DD_CHECK(ctx, DDPM3D_ERR_ARG, "finalize: null ctx");
  DD_CHECK(!ctx->finalized, DDPM3D_ERR_STATE, "finalize: already finalized");
  for (auto& p : ctx->params) DD_CHECK(p.loaded, DDPM3D_ERR_MISSING, "missing key in state_dict: " + p.key);
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    set_error("no CUDA device: this library has no CPU fallback");
    return DDPM3D_ERR_CUDA;
  }
  DD_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop{};
  DD_CUDA(cudaGetDeviceProperties(&prop, device));
  DD_CHECK(prop.major == 10, DDPM3D_ERR_CUDA, "this library is built for sm_100a (B200) only");
  ctx->device = device;
  DD_TRY(upload_f32(ctx, "time_embed.0.weight", &ctx->te_w0));
  DD_TRY(upload_f32(ctx, "time_embed.0.bias", &ctx->te_b0));
  DD_TRY(upload_f32(ctx, "time_embed.2.weight", &ctx->te_w2));
  DD_TRY(upload_f32(ctx, "time_embed.2.bias", &ctx->te_b2));
  if (ctx->cfg.num_classes > 0) DD_TRY(upload_f32(ctx, "label_emb.weight", &ctx->label_emb));
  std::vector<float> emb_w, emb_b;
  for (auto& blk : ctx->input_blocks)
    for (auto& L : blk) DD_TRY(finalize_layer(ctx, L, emb_w, emb_b));
  for (auto& L : ctx->middle) DD_TRY(finalize_layer(ctx, L, emb_w, emb_b));
  for (auto& blk : ctx->output_blocks)
    for (auto& L : blk) DD_TRY(finalize_layer(ctx, L, emb_w, emb_b));
  ctx->emb_rows_total = (int)emb_b.size();
  if (!emb_b.empty()) {
    DD_TRY(upload(ctx, emb_w.data(), emb_w.size() * sizeof(float), (void**)&ctx->emb_w_all));
    DD_TRY(upload(ctx, emb_b.data(), emb_b.size() * sizeof(float), (void**)&ctx->emb_b_all));
  }
  DD_TRY(upload_f32(ctx, "out.0.weight", &ctx->out_gn_g));
  DD_TRY(upload_f32(ctx, "out.0.bias", &ctx->out_gn_b));
  DD_TRY(pack_conv(ctx, DDPM3D_FP32, "out.2.weight", "out.2.bias", 27, nullptr, nullptr, &ctx->out_conv));
  // the first conv runs in the torso dtype
  // (input_blocks.0.0 is a plain L_CONV layer and was packed above)
  for (auto& p : ctx->params) { p.host.clear(); p.host.shrink_to_fit(); }
  DD_CUDA(cudaDeviceSynchronize());
  ctx->finalized = true;
  return DDPM3D_OK;
}

int ddpm3d_set_timestep_freqs(ddpm3d_ctx* ctx, const float* freqs, int n) {
  DD_CHECK(ctx && freqs, DDPM3D_ERR_ARG, "set_timestep_freqs: null argument");
  DD_CHECK(ctx->finalized, DDPM3D_ERR_STATE, "set_timestep_freqs: finalize weights first");
  DD_CHECK(n == ctx->cfg.model_channels / 2, DDPM3D_ERR_ARG, "set_timestep_freqs: need model_channels/2 frequencies");
  DD_CUDA(cudaSetDevice(ctx->device));
  DD_CUDA(cudaDeviceSynchronize());
  if (ctx->d_freqs) cudaFree(ctx->d_freqs);
  ctx->d_freqs = nullptr;
  DD_CUDA(cudaMalloc((void**)&ctx->d_freqs, (size_t)n * sizeof(float)));
  DD_CUDA(cudaMemcpy(ctx->d_freqs, freqs, (size_t)n * sizeof(float), cudaMemcpyDefault));
  ctx->n_freqs = n;
  return DDPM3D_OK;
}

int64_t ddpm3d_workspace_bytes(ddpm3d_ctx* ctx, int B, int Z, int H, int W) {
  if (!ctx) { set_error("workspace_bytes: null ctx"); return DDPM3D_ERR_ARG; }
  const int f = 1 << (ctx->cfg.n_levels - 1);
  if (B < 1 || Z < 1 || H < 1 || W < 1 || H % f || W % f) { set_error("workspace_bytes: bad geometry"); return DDPM3D_ERR_ARG; }
  return dry_bytes(ctx, B, Z, H, W, nullptr);
}

int ddpm3d_unet_forward(ddpm3d_ctx* ctx, const float* x, const float* low_res, const float* t, const int64_t* y, float* out,
                        int B, int Z, int H, int W, void* stream) {
  DD_CHECK(ctx && x && t && out, DDPM3D_ERR_ARG, "unet_forward: null argument");
  DD_CHECK((low_res != nullptr) == (ctx->cfg.unconditional == 0), DDPM3D_ERR_ARG,
           "unet_forward: low_res must be given to a conditional model and only to it");
  DD_TRY(check_geometry(ctx, B, Z, H, W));
  DD_CUDA(cudaSetDevice(ctx->device));
  DD_TRY(ensure_workspace(ctx, B, Z, H, W));
  cudaStream_t s = (cudaStream_t)stream;
  DD_TRY(sync_peers(ctx, s));
  GraphKey key{};
  key.kind = 0; key.B = B; key.Z = Z; key.H = H; key.W = W;
  key.p[0] = x; key.p[1] = low_res; key.p[2] = t; key.p[3] = y; key.p[4] = out; key.p[5] = s;
  return run_graphed(ctx, key, s, [&](cudaStream_t cs, int* n) { return forward_launch(ctx, x, low_res, t, y, out, B, Z, H, W, cs, n); });
}

int ddpm3d_set_schedule(ddpm3d_ctx* ctx, const ddpm3d_step_scalars* table, int T, int mean_type, int var_type) {
  DD_CHECK(ctx && table && T >= 1, DDPM3D_ERR_ARG, "set_schedule: bad argument");
  DD_CHECK(mean_type >= 0 && mean_type <= 2 && var_type >= 0 && var_type <= 3, DDPM3D_ERR_ARG, "set_schedule: bad mode");
  if (ctx->device < 0) {  // sampler-only context (no weights): bind to the caller's current device
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
      cudaGetLastError();
      set_error("no CUDA device: this library has no CPU fallback");
      return DDPM3D_ERR_CUDA;
    }
    int d = 0;
    DD_CUDA(cudaGetDevice(&d));
    ctx->device = d;
  }
  DD_CUDA(cudaSetDevice(ctx->device));
  DD_CUDA(cudaDeviceSynchronize());
  for (auto& kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
  ctx->graphs.clear();
  ctx->graph_launches.clear();
  if (ctx->d_table) cudaFree(ctx->d_table);
  ctx->d_table = nullptr;
  DD_CUDA(cudaMalloc((void**)&ctx->d_table, (size_t)T * sizeof(ddpm3d_step_scalars)));
  DD_CUDA(cudaMemcpy(ctx->d_table, table, (size_t)T * sizeof(ddpm3d_step_scalars), cudaMemcpyHostToDevice));
  ctx->T = T;
  ctx->mean_type = mean_type;
  ctx->var_type = var_type;
  return DDPM3D_OK;
}

int ddpm3d_set_sampler(ddpm3d_ctx* ctx, int kind, float eta) {
  DD_CHECK(ctx && (kind == 0 || kind == 1), DDPM3D_ERR_ARG, "set_sampler: kind must be 0 (DDPM) or 1 (DDIM)");
  if (ctx->ddim == kind && ctx->eta == eta) return DDPM3D_OK;
  if (ctx->device >= 0) {  // cached graphs bake the sampler in
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto& kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
    ctx->graphs.clear();
    ctx->graph_launches.clear();
  }
  ctx->ddim = kind;
  ctx->eta = eta;
  return DDPM3D_OK;
}

int ddpm3d_p_sample_update(ddpm3d_ctx* ctx, const float* x, const float* model_out, const float* noise, const int32_t* t_index,
                           int clip_denoised, float* sample, float* pred_xstart, float* mean, float* log_variance, int B, int C,
                           int64_t n_spatial, void* stream) {
  DD_CHECK(ctx && x && model_out && noise && t_index && sample, DDPM3D_ERR_ARG, "p_sample_update: null argument");
  DD_CHECK(ctx->d_table, DDPM3D_ERR_STATE, "p_sample_update: schedule not set");
  DD_CUDA(cudaSetDevice(ctx->device));
  UpdateArgs a{};
  a.x = x; a.model_out = model_out; a.noise = noise; a.t_index = t_index; a.table = ctx->d_table;
  a.mean_type = ctx->mean_type; a.var_type = ctx->var_type; a.clip = clip_denoised;
    a.ddim = ctx->ddim; a.eta = ctx->eta;
  a.sample = sample; a.pred_xstart = pred_xstart; a.mean = mean; a.log_variance = log_variance;
  a.B = B; a.C = C; a.n = n_spatial; a.T = ctx->T;
  DD_TRY(p_sample_update_k(a, (cudaStream_t)stream));
  ctx->launches += 1;
  return DDPM3D_OK;
}

int ddpm3d_p_sample(ddpm3d_ctx* ctx, const float* x, const float* low_res, const int64_t* y, const float* noise, int step_index,
                    int clip_denoised, float* sample, float* pred_xstart, int B, int Z, int H, int W, void* stream) {
  DD_CHECK(ctx && x && low_res && noise && sample, DDPM3D_ERR_ARG, "p_sample: null argument");
  DD_TRY(check_geometry(ctx, B, Z, H, W));
  DD_TRY(check_sampler_shapes(ctx));
  DD_CHECK(ctx->d_table && step_index >= 0 && step_index < ctx->T, DDPM3D_ERR_ARG, "p_sample: step index out of range");
  DD_CUDA(cudaSetDevice(ctx->device));
  const int64_t n = (int64_t)Z * H * W;
  DD_TRY(ensure_workspace(ctx, B, Z, H, W));
  DD_TRY(ensure_loop_buffers(ctx, B, n));
  cudaStream_t s = (cudaStream_t)stream;
  DD_TRY(sync_peers(ctx, s));
  // the step index is data (device counter), so one graph serves every step
  DD_TRY(step_set_k(ctx->d_counter, ctx->d_tmodel, ctx->d_table, B, step_index, 0, s));
  ctx->launches += 1;
  GraphKey key{};
  key.kind = 1; key.B = B; key.Z = Z; key.H = H; key.W = W;
  key.p[0] = x; key.p[1] = low_res; key.p[2] = y; key.p[3] = noise; key.p[4] = sample; key.p[5] = pred_xstart; key.p[6] = s;
  key.i[0] = clip_denoised;
  return run_graphed(ctx, key, s, [&](cudaStream_t cs, int* nl) {
    DD_TRY(forward_launch(ctx, x, low_res, ctx->d_tmodel, y, ctx->d_mo, B, Z, H, W, cs, nl));
    UpdateArgs a{};
    a.x = x; a.model_out = ctx->d_mo; a.noise = noise; a.step_counter = ctx->d_counter; a.table = ctx->d_table;
    a.mean_type = ctx->mean_type; a.var_type = ctx->var_type; a.clip = clip_denoised;
    a.ddim = ctx->ddim; a.eta = ctx->eta;
    a.sample = sample; a.pred_xstart = pred_xstart;
    a.B = B; a.C = 1; a.n = n; a.T = ctx->T;
    DD_TRY(p_sample_update_k(a, cs));
    *nl += 1;
    return (int)DDPM3D_OK;
  });
}

int ddpm3d_p_sample_t(ddpm3d_ctx* ctx, const float* x, const float* low_res, const int64_t* y, const float* noise,
                      const int64_t* t, int clip_denoised, float* sample, float* pred_xstart, int B, int Z, int H, int W,
                      void* stream) {
  DD_CHECK(ctx && x && low_res && noise && sample && t, DDPM3D_ERR_ARG, "p_sample_t: null argument");
  DD_TRY(check_geometry(ctx, B, Z, H, W));
  DD_TRY(check_sampler_shapes(ctx));
  DD_CUDA(cudaSetDevice(ctx->device));
  const int64_t n = (int64_t)Z * H * W;
  DD_TRY(ensure_workspace(ctx, B, Z, H, W));
  DD_TRY(ensure_loop_buffers(ctx, B, n));
  cudaStream_t s = (cudaStream_t)stream;
  DD_TRY(sync_peers(ctx, s));
  GraphKey key{};
  key.kind = 3; key.B = B; key.Z = Z; key.H = H; key.W = W;
  key.p[0] = x; key.p[1] = low_res; key.p[2] = y; key.p[3] = noise; key.p[4] = sample; key.p[5] = pred_xstart; key.p[6] = s;
  key.p[7] = t;
  key.i[0] = clip_denoised;
  return run_graphed(ctx, key, s, [&](cudaStream_t cs, int* nl) {
    DD_TRY(step_from_tensor_k(t, ctx->d_tindex, ctx->d_tmodel, ctx->d_table, B, ctx->T, cs));
    DD_TRY(forward_launch(ctx, x, low_res, ctx->d_tmodel, y, ctx->d_mo, B, Z, H, W, cs, nl));
    UpdateArgs a{};
    a.x = x; a.model_out = ctx->d_mo; a.noise = noise; a.t_index = ctx->d_tindex; a.table = ctx->d_table;
    a.mean_type = ctx->mean_type; a.var_type = ctx->var_type; a.clip = clip_denoised;
    a.ddim = ctx->ddim; a.eta = ctx->eta;
    a.sample = sample; a.pred_xstart = pred_xstart;
    a.B = B; a.C = 1; a.n = n; a.T = ctx->T;
    DD_TRY(p_sample_update_k(a, cs));
    *nl += 2;
    return (int)DDPM3D_OK;
  });
}

int ddpm3d_sample_loop(ddpm3d_ctx* ctx, const float* x_T, const float* low_res, const int64_t* y, const float* noise,
                       uint64_t seed, int clip_denoised, int n_steps, float* out, int B, int Z, int H, int W, void* stream) {
  DD_CHECK(ctx && x_T && low_res && out, DDPM3D_ERR_ARG, "sample_loop: null argument");
  DD_TRY(check_geometry(ctx, B, Z, H, W));
  DD_TRY(check_sampler_shapes(ctx));
  DD_CUDA(cudaSetDevice(ctx->device));
  const int64_t n = (int64_t)Z * H * W;
  DD_TRY(ensure_workspace(ctx, B, Z, H, W));
  DD_TRY(ensure_loop_buffers(ctx, B, n));
  if (n_steps <= 0 || n_steps > ctx->T) n_steps = ctx->T;
  cudaStream_t s = (cudaStream_t)stream;
  DD_TRY(sync_peers(ctx, s));
  float* img = ctx->d_img;
  DD_CUDA(cudaMemcpyAsync(img, x_T, (size_t)B * n * sizeof(float), cudaMemcpyDeviceToDevice, s));
  DD_TRY(step_set_k(ctx->d_counter, ctx->d_tmodel, ctx->d_table, B, ctx->T - 1, 0, s));
  ctx->launches += 1;
  GraphKey key{};
  key.kind = 2; key.B = B; key.Z = Z; key.H = H; key.W = W;
  key.p[0] = low_res; key.p[1] = y; key.p[2] = noise; key.p[3] = s;
  key.i[0] = clip_denoised; key.i[1] = (int64_t)seed;
  for (int k = 0; k < n_steps; ++k) {
    DD_TRY(run_graphed(ctx, key, s, [&](cudaStream_t cs, int* nl) {
      DD_TRY(forward_launch(ctx, img, low_res, ctx->d_tmodel, y, ctx->d_mo, B, Z, H, W, cs, nl));
      UpdateArgs a{};
      a.x = img; a.model_out = ctx->d_mo; a.noise = noise; a.step_counter = ctx->d_counter; a.table = ctx->d_table;
      a.mean_type = ctx->mean_type; a.var_type = ctx->var_type; a.clip = clip_denoised;
    a.ddim = ctx->ddim; a.eta = ctx->eta;
      a.sample = img;  // in place: purely elementwise
      a.B = B; a.C = 1; a.n = n; a.T = ctx->T;
      a.noise_step_stride = (int64_t)B * n;
      a.use_philox = noise == nullptr; a.seed = seed;
      if (ctx->slab.active()) {  // one noise field for the whole volume, each slab draws its part
        a.idx_bstride = (int64_t)ctx->slab.z_total * H * W;
        a.idx_offset = (int64_t)ctx->slab.z_begin * H * W;
      }
      DD_TRY(p_sample_update_k(a, cs));
      DD_TRY(step_advance_k(ctx->d_counter, ctx->d_tmodel, ctx->d_table, B, cs));
      *nl += 2;
      return (int)DDPM3D_OK;
    }));
  }
  DD_CUDA(cudaMemcpyAsync(out, img, (size_t)B * n * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return DDPM3D_OK;
}

int ddpm3d_comm_unique_id(void* out128) {
  DD_CHECK(out128, DDPM3D_ERR_ARG, "comm_unique_id: null argument");
  return comm_unique_id(out128);
}

int ddpm3d_set_comm(ddpm3d_ctx* ctx, const void* id128, int rank, int world) {
  DD_CHECK(ctx && id128, DDPM3D_ERR_ARG, "set_comm: null argument");
  DD_CHECK(ctx->finalized, DDPM3D_ERR_STATE, "set_comm: finalize weights first");
  DD_CUDA(cudaSetDevice(ctx->device));
  DD_CUDA(cudaDeviceSynchronize());
  for (auto& kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
  ctx->graphs.clear();
  ctx->graph_launches.clear();
  DD_TRY(comm_init(&ctx->slab, id128, rank, world));
  // mailboxes for the peer path (CUDA IPC over NVLink); without peer access between the devices the NCCL path stays
  if (world > 1 && ctx->slab_p2p) DD_TRY(comm_peer_init(&ctx->slab));
  return DDPM3D_OK;
}

int ddpm3d_set_slab(ddpm3d_ctx* ctx, int z_begin, int z_total) {
  DD_CHECK(ctx, DDPM3D_ERR_ARG, "set_slab: null ctx");
  const bool was = ctx->slab.active();
  if (z_total == 0) {  // back to un-sharded work (independent patches, ensemble samples): no halo / statistics exchange
    DD_CHECK(z_begin == 0, DDPM3D_ERR_ARG, "set_slab: bad bounds");
    ctx->slab.enabled = false;
  } else {
    DD_CHECK(z_begin >= 0 && z_total >= 1 && z_begin < z_total, DDPM3D_ERR_ARG, "set_slab: bad bounds");
    DD_CHECK(ctx->slab.comm != nullptr, DDPM3D_ERR_STATE, "set_slab: no communicator (call ddpm3d_set_comm first)");
    ctx->slab.z_begin = z_begin;
    ctx->slab.z_total = z_total;
    ctx->slab.enabled = true;
  }
  if (was != ctx->slab.active() && ctx->device >= 0) {  // workspace layout (halo planes) and graphs depend on the mode
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto& kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
    ctx->graphs.clear();
    ctx->graph_launches.clear();
  }
  return DDPM3D_OK;
}

int ddpm3d_set_option(ddpm3d_ctx* ctx, const char* name, int64_t value) {
  DD_CHECK(ctx && name, DDPM3D_ERR_ARG, "set_option: null argument");
  const std::string n(name);
  if (n == "cuda_graph") ctx->use_graph = value != 0;
  else if (n == "conv_path") { DD_CHECK(value >= 0 && value <= 2, DDPM3D_ERR_ARG, "conv_path must be 0, 1 or 2"); ctx->conv_path = (int)value; }
  else if (n == "profile") { DD_CHECK(value >= 0 && value <= 2, DDPM3D_ERR_ARG, "profile must be 0, 1 or 2"); ctx->profile = (int)value; }
  else if (n == "fuse_stats") ctx->fuse_stats = value != 0;
  else if (n == "split_k") ctx->split_k = value != 0;
  else if (n == "cluster") ctx->cluster = value != 0;
  else if (n == "strip") ctx->strip = value < 0 ? 0 : (value > 2 ? 2 : value);
  else if (n == "strip_w") ctx->strip_w = value;
  else if (n == "fold_identity") ctx->fold_identity = value != 0;
  else if (n == "stem_tc") ctx->stem_tc = value != 0;
  else if (n == "head_v2") ctx->head_v2 = value != 0;
  else if (n == "head_tc") ctx->head_tc = value != 0;
  else if (n == "slab_p2p") ctx->slab_p2p = value != 0;
  else if (n == "pdl") ctx->pdl = value < 0 ? 0 : value;
  else if (n == "gn_stream") ctx->gn_stream = value != 0;
  else if (n == "gn_stream_mb") ctx->gn_stream_mb = (int)value;
  else { set_error("unknown option: " + n); return DDPM3D_ERR_ARG; }
  // cached graphs bake the options in
  if (ctx->device >= 0) {
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto& kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
    ctx->graphs.clear();
    ctx->graph_launches.clear();
  }
  return DDPM3D_OK;
}

int64_t ddpm3d_launch_count(const ddpm3d_ctx* ctx) { return ctx ? ctx->launches : 0; }

int ddpm3d_profile_read(ddpm3d_ctx* ctx, ddpm3d_prof_record* out, int cap) {
  DD_CHECK(ctx, DDPM3D_ERR_ARG, "profile_read: null ctx");
  if (ctx->device >= 0) {
    DD_CUDA(cudaSetDevice(ctx->device));
    DD_CUDA(cudaDeviceSynchronize());
  }
  const int n = (int)ctx->prof.size();
  for (int i = 0; i < n; ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->prof[i].a, ctx->prof[i].b);
    if (out && i < cap) {
      out[i].kind = ctx->prof[i].kind;
      out[i].pad_ = 0;
      out[i].ms = ms;
      out[i].pad2_ = 0.f;
      out[i].work = ctx->prof[i].work;
    }
    cudaEventDestroy(ctx->prof[i].a);
    cudaEventDestroy(ctx->prof[i].b);
  }
  ctx->prof.clear();
  if (ctx->profile == 2) {  // the captured graphs reference the events just destroyed
    for (auto& kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
    ctx->graphs.clear();
    ctx->graph_launches.clear();
  }
  return n;
}

// ---- single kernels (unit tests) ------------------------------------------------------------------
int ddpm3d_k_conv3d(int dtype, int path, const void* in, const void* w, const float* bias, const void* residual, void* out,
                    int B, int Z, int H, int W, int Cin, int Cout, int taps, int stride_hw, void* stream) {
  DD_CHECK(in && w && out, DDPM3D_ERR_ARG, "k_conv3d: null argument");
  DD_CHECK(stride_hw == 1 || stride_hw == 2, DDPM3D_ERR_ARG, "k_conv3d: stride_hw must be 1 or 2");
  DD_CHECK(H % stride_hw == 0 && W % stride_hw == 0, DDPM3D_ERR_ARG, "k_conv3d: H, W must be divisible by the stride");
  ConvArgs a{};
  a.dt = dtype; a.main = {in, Cin}; a.taps = taps; a.stride_hw = stride_hw; a.w = w; a.bias = bias;
  a.residual = residual; a.res_mode = residual ? RES_SAME : RES_NONE;
  a.out = out; a.B = B; a.Z = Z; a.Ho = H / stride_hw; a.Wo = W / stride_hw; a.Cout = Cout;
  if (path == 2 || (path >= 5 && path <= 7)) {
    DD_CHECK(is_half_dt(dtype) && conv_tc_eligible(a), DDPM3D_ERR_ARG, "k_conv3d: shape not eligible for the tcgen05 path");
    if (path == 5 || path == 6) a.strip_allowed = path == 5 ? 0 : 1;
    if (path == 7) a.strip_maxw = 4;
    a.splitk_allowed = 1;
    const size_t need = conv_tc_scratch_bytes(a);
    if (need) {
      void* sc = nullptr;
      DD_TRY(g_scratch.get(need, &sc));
      a.splitk_scratch = (float*)sc;
      a.splitk_bytes = need;
    }
    return conv_tc(a, (cudaStream_t)stream);
  }
  if (path == 3 || path == 4) {  // the Cin == 2 stem: 3 = tensor-core tile per 128 voxels (16-bit), 4 = CUDA cores
    DD_CHECK(conv_stem_eligible(a), DDPM3D_ERR_ARG, "k_conv3d: shape not eligible for the stem kernels");
    a.stem_tc_allowed = path == 3;
    return conv_stem(a, (cudaStream_t)stream);
  }
  return conv_simt(a, (cudaStream_t)stream);
}

int ddpm3d_k_conv_plan(int dtype, int B, int Z, int H, int W, int Cin, int Cout, int taps, int extra_channels, int split_k,
                       int strip, int sms, int32_t* out8) {
  DD_CHECK(out8 != nullptr && B >= 1 && Z >= 1 && H >= 1 && W >= 1, DDPM3D_ERR_ARG, "k_conv_plan: bad argument");
  ConvArgs a{};
  a.dt = dtype; a.main = {nullptr, Cin}; a.taps = taps; a.B = B; a.Z = Z; a.Ho = H; a.Wo = W; a.Cout = Cout;
  if (extra_channels > 0) { a.n_extra = 1; a.extra[0] = {nullptr, extra_channels}; }
  a.splitk_allowed = split_k;
  a.strip_allowed = strip;
  int v[8];
  DD_TRY(conv_tc_plan_query(a, sms, v));
  for (int i = 0; i < 8; ++i) out8[i] = v[i];
  return DDPM3D_OK;
}

int ddpm3d_k_probe_rowshift(const void* a, int rows, const void* ident, int shift, int mode, float* out, void* stream) {
  DD_CHECK(a && ident && out, DDPM3D_ERR_ARG, "k_probe_rowshift: null argument");
  return probe_rowshift(a, rows, ident, shift, mode, out, (cudaStream_t)stream);
}

int ddpm3d_k_groupnorm(int dtype, const void* in, const float* gamma, const float* beta, const float* film, int silu,
                       int resample, void* out, int B, int Z, int H, int W, int C, void* stream) {
  DD_CHECK(in && gamma && beta && out, DDPM3D_ERR_ARG, "k_groupnorm: null argument");
  GnArgs g{};
  g.dt = dtype; g.src[0] = in; g.C[0] = C; g.B = B; g.Z = Z; g.H = H; g.W = W; g.gamma = gamma; g.beta = beta;
  g.film = film; g.film_stride = 2 * (int64_t)C; g.silu = silu; g.resample = resample; g.out = out;
  g.n_chunks = gn_chunks((int64_t)Z * H * W);
  const size_t pb = (size_t)B * g.n_chunks * 64 * sizeof(double), ab = (size_t)B * 2 * C * sizeof(float);
  void* scratch = nullptr;
  DD_TRY(g_scratch.get(pb + ab + 256, &scratch));
  g.partials = (double*)scratch;
  g.ab = (float*)((char*)scratch + ((pb + 255) & ~size_t(255)));
  int n = 0;
  return gn_forward(g, (cudaStream_t)stream, &n);
}

int ddpm3d_k_conv3d_gn(int dtype, const void* in, const void* w, const float* bias, const float* gamma, const float* beta,
                       void* conv_out, void* gn_out, int B, int Z, int H, int W, int Cin, int Cout, void* stream) {
  DD_CHECK(in && w && bias && gamma && beta && conv_out && gn_out, DDPM3D_ERR_ARG, "k_conv3d_gn: null argument");
  ConvArgs a{};
  a.dt = dtype; a.main = {in, Cin}; a.taps = 27; a.w = w; a.bias = bias;
  a.out = conv_out; a.B = B; a.Z = Z; a.Ho = H; a.Wo = W; a.Cout = Cout;
  DD_CHECK(is_half_dt(dtype) && conv_tc_eligible(a), DDPM3D_ERR_ARG, "k_conv3d_gn: shape not eligible for the tcgen05 path");
  GnArgs g{};
  g.dt = dtype; g.src[0] = conv_out; g.C[0] = Cout; g.B = B; g.Z = Z; g.H = H; g.W = W; g.gamma = gamma; g.beta = beta;
  g.silu = 0; g.out = gn_out;
  const size_t cs = (size_t)B * chsum_slots() * Cout * 2 * sizeof(float), ab = (size_t)B * 2 * Cout * sizeof(float);
  void* scratch = nullptr;
  DD_TRY(g_scratch.get(cs + ab + 512, &scratch));
  a.chsum_out = (float*)scratch;
  g.ab = (float*)((char*)scratch + ((cs + 255) & ~size_t(255)));
  a.splitk_allowed = 0;  // stream-K layers do not produce channel sums
  DD_TRY(conv_tc(a, (cudaStream_t)stream));
  DD_CHECK(a.chsum_written, DDPM3D_ERR_STATE, "k_conv3d_gn: this shape does not produce epilogue channel sums");
  g.chsum[0] = a.chsum_out;
  g.chsum_bias[0] = bias;
  return gn_forward_chsum(g, (cudaStream_t)stream);
}

int ddpm3d_k_timestep_embedding(const float* t, const float* freqs, float* out, int B, int dim, void* stream) {
  DD_CHECK(t && out && B >= 1 && dim >= 1, DDPM3D_ERR_ARG, "k_timestep_embedding: bad argument");
  return timestep_embedding_k(t, freqs, out, B, dim, (cudaStream_t)stream);
}

int ddpm3d_k_attention(int dtype, const void* qkv, void* out, int B, int T, int C, int heads, int new_order, void* stream) {
  DD_CHECK(qkv && out, DDPM3D_ERR_ARG, "k_attention: null argument");
  const int path = new_order >> 8;  // bits 8..: 1 forces the CUDA-core kernel (unit tests cross-check the two)
  new_order &= 0xff;
  const size_t need = path == 1 ? 0 : attention_tc_scratch_bytes(dtype, B, T, C, heads);
  void* sc = nullptr;
  if (need) DD_TRY(g_scratch.get(need, &sc));
  return attention_k(dtype, qkv, out, B, T, C, heads, new_order, sc, need, (cudaStream_t)stream);
}

int ddpm3d_k_attention_window(int dtype, const void* qkv, void* out, int B, int T, int C, int heads, int new_order,
                              int q_begin, int q_count, void* stream) {
  DD_CHECK(qkv && out, DDPM3D_ERR_ARG, "k_attention_window: null argument");
  const int path = new_order >> 8;
  new_order &= 0xff;
  const size_t need = path == 1 ? 0 : attention_tc_scratch_bytes(dtype, B, T, C, heads);
  void* sc = nullptr;
  if (need) DD_TRY(g_scratch.get(need, &sc));
  return attention_k(dtype, qkv, out, B, T, C, heads, new_order, sc, need, (cudaStream_t)stream, q_begin, q_count);
}

}  // extern "C"
