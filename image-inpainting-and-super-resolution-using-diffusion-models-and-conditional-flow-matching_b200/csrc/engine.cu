// Sampling engine: plan builder, U-Net executor, samplers and the C ABI (include/cfm_b200.h).
//
// The U-Net of AD/image_diffusion/unet.py:490-728 is lowered once, at cfm_engine_create, into
// a flat list of ops over NHWC activation tensors that live in one arena (offsets assigned by
// a liveness-based first-fit allocator).  An NFE is then a fixed sequence of kernel launches
// with no host synchronisation, which is what makes whole-loop CUDA-graph capture possible.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include "engine.h"
#include "kernels_generic.cuh"
#include "steps.cuh"

namespace cfm {

static thread_local std::string g_create_error = "";

#define CU_CHECK(e, call)                                                                     \
  do {                                                                                        \
    cudaError_t _st = (call);                                                                 \
    if (_st != cudaSuccess) {                                                                 \
      (e).err = std::string(#call) + ": " + cudaGetErrorString(_st);                          \
      return CFM_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

static int fail(Engine& e, int code, const std::string& msg) { e.err = msg; return code; }

// ---------------------------------------------------------------------------------------------
// weights
// ---------------------------------------------------------------------------------------------
static int fetch(Engine& e, const std::string& name, int64_t numel, const float** out) {
  auto it = e.sd.find(name);
  if (it == e.sd.end()) return fail(e, CFM_ERR_MISSING, "state_dict is missing '" + name + "'");
  if (it->second.numel != numel)
    return fail(e, CFM_ERR_MISSING, "state_dict tensor '" + name + "' has " + std::to_string(it->second.numel) +
                                        " elements, expected " + std::to_string(numel));
  *out = it->second.data;
  e.param_count += numel;
  return 0;
}

static int upload(Engine& e, const float* host, size_t n, float** dev) {
  void* p = nullptr;
  CU_CHECK(e, cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(float)));
  e.owned.push_back(p);
  CU_CHECK(e, cudaMemcpy(p, host, n * sizeof(float), cudaMemcpyHostToDevice));
  *dev = (float*)p;
  return 0;
}

template <typename T>
static int dev_alloc(Engine& e, size_t n, T** dev) {
  void* p = nullptr;
  CU_CHECK(e, cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
  e.owned.push_back(p);
  *dev = (T*)p;
  return 0;
}

// OIHW [Cout][Cin][ks][ks] -> [(ky*ks+kx)*Cin + ci][Cout]
static std::vector<float> to_kn(const float* w, int Cout, int Cin, int ks) {
  std::vector<float> r((size_t)ks * ks * Cin * Cout);
  for (int o = 0; o < Cout; ++o)
    for (int c = 0; c < Cin; ++c)
      for (int t = 0; t < ks * ks; ++t)
        r[((size_t)t * Cin + c) * Cout + o] = w[((size_t)o * Cin + c) * ks * ks + t];
  return r;
}

// ---------------------------------------------------------------------------------------------
// plan construction
// ---------------------------------------------------------------------------------------------
struct Cur { int id; int C, H, W; };

static int new_tensor(Engine& e, int C, int H, int W) {
  TensorDesc t; t.C = C; t.H = H; t.W = W;
  e.tensors.push_back(t);
  return (int)e.tensors.size() - 1;
}

static int heads_for(const cfm_unet_config& c, int channels, bool upsample_side) {
  if (c.num_head_channels != -1) return channels / c.num_head_channels;
  if (upsample_side && c.num_heads_upsample != -1) return c.num_heads_upsample;
  return c.num_heads;
}

// Adds a conv op; w given as state_dict names.  skip_name empty => no 1x1 skip operand.
static int add_conv(Engine& e, const std::string& name, const std::string& wname, int ks, int stride, int ups,
                    int src0, int src1, bool src_is_input, int Cin, int Hin, int Win, int Cout,
                    const std::string& skip_wname, int skip0, int skip1, int Cskip,
                    int res0, int res1, int emb_off, int out, bool out_is_output, const std::string& gn_name = "") {
  Op op; op.kind = OP_CONV; op.name = name;
  op.ks = ks; op.stride = stride; op.ups = ups; op.src0 = src0; op.src1 = src1; op.src_is_input = src_is_input;
  op.Cin = Cin; op.Hin = Hin; op.Win = Win; op.Cout = Cout;
  const int Hv = ups ? Hin * 2 : Hin, Wv = ups ? Win * 2 : Win;
  op.Hout = (Hv + 2 * (ks / 2) - ks) / stride + 1; op.Wout = (Wv + 2 * (ks / 2) - ks) / stride + 1;
  op.skip0 = skip0; op.skip1 = skip1; op.Cskip = Cskip; op.res0 = res0; op.res1 = res1;
  op.emb_off = emb_off; op.out = out; op.out_is_output = out_is_output;
  const float *w = nullptr, *b = nullptr, *ws = nullptr, *bs = nullptr;
  int rc;
  if ((rc = fetch(e, wname + ".weight", (int64_t)Cout * Cin * ks * ks, &w))) return rc;
  if ((rc = fetch(e, wname + ".bias", Cout, &b))) return rc;
  std::vector<float> bias(b, b + Cout);
  std::vector<float> wkn = to_kn(w, Cout, Cin, ks);
  if ((rc = upload(e, wkn.data(), wkn.size(), &op.w_main))) return rc;
  std::vector<float> w_oihw(w, w + (size_t)Cout * Cin * ks * ks), ws_oi;
  if (!skip_wname.empty()) {
    if ((rc = fetch(e, skip_wname + ".weight", (int64_t)Cout * Cskip, &ws))) return rc;
    if ((rc = fetch(e, skip_wname + ".bias", Cout, &bs))) return rc;
    for (int i = 0; i < Cout; ++i) bias[i] += bs[i];
    std::vector<float> wskn = to_kn(ws, Cout, Cskip, 1);
    if ((rc = upload(e, wskn.data(), wskn.size(), &op.w_skip))) return rc;
    ws_oi.assign(ws, ws + (size_t)Cout * Cskip);
  }
  if ((rc = upload(e, bias.data(), bias.size(), &op.bias))) return rc;
  op.flops = 2.0 * op.Hout * op.Wout * Cout * ((double)ks * ks * Cin + Cskip);
  if (e.bf16 && tc_conv_supported(e, op)) {
    if (!gn_name.empty()) {     // ask the tensor-core kernel to apply the GroupNorm (+SiLU) that follows in its epilogue
      const int64_t before = e.param_count;
      const float *g = nullptr, *b2 = nullptr;
      if ((rc = fetch(e, gn_name + ".weight", Cout, &g))) return rc;
      if ((rc = fetch(e, gn_name + ".bias", Cout, &b2))) return rc;
      if ((rc = upload(e, g, Cout, &op.gamma))) return rc;
      if ((rc = upload(e, b2, Cout, &op.beta))) return rc;
      op.gn_request = true; op.silu = 1;
      e.param_count = before;   // counted when the GroupNorm is really folded in (below) or by add_gn
    }
    if ((rc = tc_conv_prepare(e, op, w_oihw, ws_oi))) return rc;
    e.n_tc_convs++;
    if (op.gn_fused) {
      e.param_count += 2 * Cout;
      op.name += "+" + gn_name.substr(gn_name.rfind("out_layers") == std::string::npos ? 0 : gn_name.rfind("out_layers"));
      if (op.gn_ctas > 1) { op.gn_exch_off = e.gn_tiles_per_sample; e.gn_tiles_per_sample += op.gn_ctas; }
    }
  }
  e.ops.push_back(op);
  return 0;
}

static int add_gn(Engine& e, const std::string& name, const std::string& pname, int src0, int src1, int C,
                  int silu, bool film, int emb_off, int out) {
  Op op; op.kind = OP_GN; op.name = name; op.src0 = src0; op.src1 = src1; op.out = out;
  op.Cin = C; op.silu = silu; op.film = film; op.emb_off = emb_off;
  op.Hin = e.tensors[out].H; op.Win = e.tensors[out].W;
  if (C % 32) return fail(e, CFM_ERR_INVALID, "GroupNorm32 needs channels divisible by 32 at " + name);
  const float *g = nullptr, *b = nullptr; int rc;
  if ((rc = fetch(e, pname + ".weight", C, &g))) return rc;
  if ((rc = fetch(e, pname + ".bias", C, &b))) return rc;
  if ((rc = upload(e, g, C, &op.gamma))) return rc;
  if ((rc = upload(e, b, C, &op.beta))) return rc;
  e.ops.push_back(op);
  return 0;
}

static void add_resample(Engine& e, const std::string& name, int src, int out, int up) {
  Op op; op.kind = OP_RESAMPLE; op.name = name; op.src0 = src; op.out = out; op.up = up;
  op.Cin = e.tensors[src].C; op.Hin = e.tensors[src].H; op.Win = e.tensors[src].W;
  e.ops.push_back(op);
}

struct EmbPiece { std::string prefix; int width; };

// ResBlock (unet.py:243-351).  x may be a concat (xa, xb); returns output tensor in *out.
static int add_resblock(Engine& e, const std::string& p, Cur xa, Cur xb, int cout, bool up, bool down,
                        std::vector<EmbPiece>& emb_pieces, Cur* outc) {
  const bool film = e.cfg.use_scale_shift_norm != 0;
  const int cin = xa.C + (xb.id >= 0 ? xb.C : 0);
  int H = xa.H, W = xa.W, rc;
  const int a1 = new_tensor(e, cin, H, W);
  if ((rc = add_gn(e, p + ".in_layers.0", p + ".in_layers.0", xa.id, xb.id, cin, 1, false, -1, a1))) return rc;
  int conv_src = a1, xs0 = xa.id, xs1 = xb.id;
  if (up || down) {
    if (xb.id >= 0) return fail(e, CFM_ERR_INVALID, "up/down ResBlock with a concatenated input is not supported");
    const int Hn = up ? H * 2 : H / 2, Wn = up ? W * 2 : W / 2;
    const int a1r = new_tensor(e, cin, Hn, Wn), xr = new_tensor(e, cin, Hn, Wn);
    add_resample(e, p + ".h_upd", a1, a1r, up ? 1 : 0);
    add_resample(e, p + ".x_upd", xa.id, xr, up ? 1 : 0);
    conv_src = a1r; xs0 = xr; xs1 = -1; H = Hn; W = Wn;
  }
  const int emb_off = e.emb_total;
  const int ew = film ? 2 * cout : cout;
  emb_pieces.push_back({p + ".emb_layers.1", ew});
  e.emb_total += ew;
  const int h1 = new_tensor(e, cout, H, W);
  // out_layers.0/1 (GroupNorm + SiLU of h) is folded into this conv's epilogue when the tensor-core kernel can hold whole
  // samples' statistics (no FiLM): h1 then already IS the normalised, activated tensor and no GroupNorm op is emitted
  if ((rc = add_conv(e, p + ".in_layers.2", p + ".in_layers.2", 3, 1, 0, conv_src, -1, false, cin, H, W, cout,
                     "", -1, -1, 0, -1, -1, film ? -1 : emb_off, h1, false, film ? "" : p + ".out_layers.0"))) return rc;
  int a2 = h1;
  if (!e.ops.back().gn_fused) {
    a2 = new_tensor(e, cout, H, W);
    if ((rc = add_gn(e, p + ".out_layers.0", p + ".out_layers.0", h1, -1, cout, 1, film, film ? emb_off : -1, a2))) return rc;
  }
  const int o = new_tensor(e, cout, H, W);
  if (cin != cout) {
    if ((rc = add_conv(e, p + ".out_layers.3", p + ".out_layers.3", 3, 1, 0, a2, -1, false, cout, H, W, cout,
                       p + ".skip_connection", xs0, xs1, cin, -1, -1, -1, o, false))) return rc;
  } else {
    if ((rc = add_conv(e, p + ".out_layers.3", p + ".out_layers.3", 3, 1, 0, a2, -1, false, cout, H, W, cout,
                       "", -1, -1, 0, xs0, xs1, -1, o, false))) return rc;
  }
  *outc = {o, cout, H, W};
  return 0;
}

// AttentionBlock (unet.py:354-401)
static int add_attention(Engine& e, const std::string& p, Cur x, int heads, Cur* outc) {
  int rc;
  const int C = x.C;
  if (heads <= 0 || C % heads) return fail(e, CFM_ERR_INVALID, "bad head count at " + p);
  const int a = new_tensor(e, C, x.H, x.W);
  if ((rc = add_gn(e, p + ".norm", p + ".norm", x.id, -1, C, 0, false, -1, a))) return rc;
  const int att = new_tensor(e, C, x.H, x.W);
  Op op; op.kind = OP_ATTN; op.name = p + ".attention"; op.out = att;
  op.heads = heads; op.ch = C / heads; op.Cin = C; op.Hin = x.H; op.Win = x.W;
  op.flops = 2.0 * 2.0 * (double)(x.H * x.W) * (x.H * x.W) * C;
  {
    const int qkv = new_tensor(e, 3 * C, x.H, x.W);
    if ((rc = add_conv(e, p + ".qkv", p + ".qkv", 1, 1, 0, a, -1, false, C, x.H, x.W, 3 * C, "", -1, -1, 0, -1, -1, -1, qkv, false))) return rc;
    op.src0 = qkv;
  }
  e.ops.push_back(op);
  const int o = new_tensor(e, C, x.H, x.W);
  if ((rc = add_conv(e, p + ".proj_out", p + ".proj_out", 1, 1, 0, att, -1, false, C, x.H, x.W, C, "", -1, -1, 0, x.id, -1, -1, o, false))) return rc;
  *outc = {o, C, x.H, x.W};
  return 0;
}

// Tensor-core stem: im2col (hi/lo bf16 split of the fp32 input) + a K_pad-deep 1x1 GEMM.  *done stays false
// (and nothing is added) when the tcgen05 kernel cannot take the shape.
static int add_stem_tc(Engine& e, int Cin, int S, int Cout, int out, bool* done) {
  const int K9 = 9 * Cin, Kpad = (2 * K9 + 63) / 64 * 64;
  Op op; op.kind = OP_CONV; op.name = "input_blocks.0.0";
  op.ks = 1; op.stride = 1; op.ups = 0; op.Cin = Kpad; op.Hin = op.Win = op.Hout = op.Wout = S; op.Cout = Cout;
  op.out = out;
  TensorDesc probe; probe.C = Kpad; probe.H = S; probe.W = S;
  e.tensors.push_back(probe);
  op.src0 = (int)e.tensors.size() - 1;
  if (!tc_conv_supported(e, op)) { e.tensors.pop_back(); return 0; }
  const float *w = nullptr, *b = nullptr; int rc;
  if ((rc = fetch(e, "input_blocks.0.0.weight", (int64_t)Cout * Cin * 9, &w))) return rc;
  if ((rc = fetch(e, "input_blocks.0.0.bias", Cout, &b))) return rc;
  std::vector<float> w_eff((size_t)Cout * Kpad, 0.f);          // [Cout][Kpad]: k = tap*Cin + c, duplicated for the lo term
  for (int o = 0; o < Cout; ++o)
    for (int cc = 0; cc < Cin; ++cc)
      for (int tap = 0; tap < 9; ++tap) {
        const float v = w[((size_t)o * Cin + cc) * 9 + tap];
        w_eff[(size_t)o * Kpad + tap * Cin + cc] = v;
        w_eff[(size_t)o * Kpad + K9 + tap * Cin + cc] = v;
      }
  std::vector<float> wkn = to_kn(w_eff.data(), Cout, Kpad, 1);
  if ((rc = upload(e, wkn.data(), wkn.size(), &op.w_main))) return rc;
  if ((rc = upload(e, b, Cout, &op.bias))) return rc;
  op.flops = 2.0 * S * S * Cout * K9;
  Op col; col.kind = OP_IM2COL; col.name = "input_blocks.0.0.im2col"; col.out = op.src0;
  col.Cin = Cin; col.Cout = Kpad; col.Hin = col.Win = S;
  e.ops.push_back(col);
  if ((rc = tc_conv_prepare(e, op, w_eff, {}))) return rc;
  e.n_tc_convs++;
  e.ops.push_back(op);
  *done = true;
  return 0;
}

// Tensor-core head: the 3x3 conv to <= 3 channels as (1) a 1x1 GEMM to the 9 x Cout per-tap partial products
// (fp32, 32 per pixel) and (2) a gather that sums each pixel's 3x3 neighbourhood of them.  *done stays false (and
// nothing is added) when the shapes do not fit.
static int add_head_tc(Engine& e, int a, Cur h, int Cout, bool* done) {
  const char* off = tuning_env("CFM_DISABLE_TC_HEAD_TAPS");
  // 2 or 3 output channels fill the 32-wide row of partial products (18 / 27 of 32); with one channel (MNIST) the fp32
  // row is mostly padding and the direct 3x3 conv measured the same
  if ((off && off[0] == '1') || 9 * Cout > 32 || 9 * Cout <= 16) return 0;
  Op op; op.kind = OP_CONV; op.name = "out.2";
  op.ks = 1; op.stride = 1; op.ups = 0; op.src0 = a; op.Cin = h.C; op.Hin = op.Hout = h.H; op.Win = op.Wout = h.W; op.Cout = 32;
  op.out_f32 = true;
  if (!tc_conv_supported(e, op)) return 0;
  const float *w = nullptr, *b = nullptr; int rc;
  if ((rc = fetch(e, "out.2.weight", (int64_t)Cout * h.C * 9, &w))) return rc;
  if ((rc = fetch(e, "out.2.bias", Cout, &b))) return rc;
  std::vector<float> w_eff((size_t)32 * h.C, 0.f);      // [32][Cin]: row tap * Cout + o
  for (int tap = 0; tap < 9; ++tap)
    for (int o = 0; o < Cout; ++o)
      for (int c = 0; c < h.C; ++c) w_eff[(size_t)(tap * Cout + o) * h.C + c] = w[((size_t)o * h.C + c) * 9 + tap];
  std::vector<float> wkn = to_kn(w_eff.data(), 32, h.C, 1), zeros(32, 0.f);
  if ((rc = upload(e, wkn.data(), wkn.size(), &op.w_main))) return rc;
  if ((rc = upload(e, zeros.data(), 32, &op.bias))) return rc;
  op.out = new_tensor(e, 64, h.H, h.W);                  // 32 fp32 per pixel = 64 bf16-sized elements
  op.flops = 2.0 * h.H * h.W * Cout * 9.0 * h.C;
  if ((rc = tc_conv_prepare(e, op, w_eff, {}))) return rc;
  e.n_tc_convs++;
  e.ops.push_back(op);
  Op g; g.kind = OP_HEAD_GATHER; g.name = "out.2.gather"; g.src0 = op.out; g.Cout = Cout; g.Hin = h.H; g.Win = h.W; g.out_is_output = true;
  if ((rc = upload(e, b, Cout, &g.bias))) return rc;
  e.ops.push_back(g);
  *done = true;
  return 0;
}

static int build_plan(Engine& e) {
  const cfm_unet_config& c = e.cfg;
  const int mc = c.model_channels;
  e.ted = 4 * mc;
  int rc;
  auto is_attn = [&](int ds) { for (int i = 0; i < c.n_attention_ds; ++i) if (c.attention_ds[i] == ds) return true; return false; };
  std::vector<EmbPiece> emb_pieces;
  std::vector<Cur> hs;
  int ch = (int)(c.channel_mult[0] * mc);
  const int S = c.image_size;
  Cur h;
  {
    const int t = new_tensor(e, ch, S, S);
    bool done = false;
    if (e.bf16 && 18 * c.in_channels <= 128) {
      if ((rc = add_stem_tc(e, c.in_channels, S, ch, t, &done))) return rc;
    }
    if (!done && (rc = add_conv(e, "input_blocks.0.0", "input_blocks.0.0", 3, 1, 0, -1, -1, true, c.in_channels, S, S, ch,
                                "", -1, -1, 0, -1, -1, -1, t, false))) return rc;
    h = {t, ch, S, S};
    hs.push_back(h);
  }
  int ds = 1, idx = 1;
  for (int level = 0; level < c.n_levels; ++level) {
    const int outc = (int)(c.channel_mult[level] * mc);
    for (int r = 0; r < c.num_res_blocks; ++r, ++idx) {
      const std::string p = "input_blocks." + std::to_string(idx);
      if ((rc = add_resblock(e, p + ".0", h, Cur{-1, 0, 0, 0}, outc, false, false, emb_pieces, &h))) return rc;
      if (is_attn(ds)) if ((rc = add_attention(e, p + ".1", h, heads_for(c, h.C, false), &h))) return rc;
      hs.push_back(h);
    }
    if (level != c.n_levels - 1) {
      const std::string p = "input_blocks." + std::to_string(idx) + ".0";
      if (c.resblock_updown) {
        if ((rc = add_resblock(e, p, h, Cur{-1, 0, 0, 0}, h.C, false, true, emb_pieces, &h))) return rc;
      } else if (c.conv_resample) {
        const int t = new_tensor(e, h.C, h.H / 2, h.W / 2);   // k3 s2 p1: floor((H-1)/2)+1
        e.tensors[t].H = (h.H - 1) / 2 + 1; e.tensors[t].W = (h.W - 1) / 2 + 1;
        if ((rc = add_conv(e, p + ".op", p + ".op", 3, 2, 0, h.id, -1, false, h.C, h.H, h.W, h.C, "", -1, -1, 0, -1, -1, -1, t, false))) return rc;
        h = {t, h.C, e.tensors[t].H, e.tensors[t].W};
      } else {
        const int t = new_tensor(e, h.C, h.H / 2, h.W / 2);
        add_resample(e, p + ".op", h.id, t, 0);
        h = {t, h.C, h.H / 2, h.W / 2};
      }
      hs.push_back(h);
      ds *= 2; ++idx;
    }
  }
  if ((rc = add_resblock(e, "middle_block.0", h, Cur{-1, 0, 0, 0}, h.C, false, false, emb_pieces, &h))) return rc;
  if ((rc = add_attention(e, "middle_block.1", h, heads_for(c, h.C, false), &h))) return rc;
  if ((rc = add_resblock(e, "middle_block.2", h, Cur{-1, 0, 0, 0}, h.C, false, false, emb_pieces, &h))) return rc;
  idx = 0;
  for (int level = c.n_levels - 1; level >= 0; --level) {
    const int outc = (int)(mc * c.channel_mult[level]);
    for (int i = 0; i <= c.num_res_blocks; ++i, ++idx) {
      const std::string p = "output_blocks." + std::to_string(idx);
      const Cur skip = hs.back(); hs.pop_back();
      if (skip.H != h.H || skip.W != h.W) return fail(e, CFM_ERR_INVALID, "skip/feature size mismatch (image_size not divisible by 2^levels)");
      int sub = 0;
      if ((rc = add_resblock(e, p + "." + std::to_string(sub++), h, skip, outc, false, false, emb_pieces, &h))) return rc;
      if (is_attn(ds)) if ((rc = add_attention(e, p + "." + std::to_string(sub++), h, heads_for(c, h.C, true), &h))) return rc;
      if (level && i == c.num_res_blocks) {
        const std::string pu = p + "." + std::to_string(sub++);
        if (c.resblock_updown) {
          if ((rc = add_resblock(e, pu, h, Cur{-1, 0, 0, 0}, h.C, true, false, emb_pieces, &h))) return rc;
        } else if (c.conv_resample) {
          const int t = new_tensor(e, h.C, h.H * 2, h.W * 2);
          Op probe; probe.kind = OP_CONV; probe.ks = 3; probe.stride = 1; probe.ups = 1; probe.src0 = h.id; probe.Cin = probe.Cout = h.C;
          probe.Hin = h.H; probe.Win = h.W; probe.Hout = 2 * h.H; probe.Wout = 2 * h.W;
          if (e.bf16 && !tc_conv_supported(e, probe)) {
            // bf16 shapes the tensor-core kernel cannot fold: materialise the nearest-neighbour upsample, then a plain 3x3 conv
            const int u = new_tensor(e, h.C, h.H * 2, h.W * 2);
            add_resample(e, pu + ".interpolate", h.id, u, 1);
            if ((rc = add_conv(e, pu + ".conv", pu + ".conv", 3, 1, 0, u, -1, false, h.C, h.H * 2, h.W * 2, h.C, "", -1, -1, 0, -1, -1, -1, t, false))) return rc;
          } else {
            if ((rc = add_conv(e, pu + ".conv", pu + ".conv", 3, 1, 1, h.id, -1, false, h.C, h.H, h.W, h.C, "", -1, -1, 0, -1, -1, -1, t, false))) return rc;
          }
          h = {t, h.C, h.H * 2, h.W * 2};
        } else {
          const int t = new_tensor(e, h.C, h.H * 2, h.W * 2);
          add_resample(e, pu, h.id, t, 1);
          h = {t, h.C, h.H * 2, h.W * 2};
        }
        ds /= 2;
      }
    }
  }
  {
    const int a = new_tensor(e, h.C, h.H, h.W);
    if ((rc = add_gn(e, "out.0", "out.0", h.id, -1, h.C, 1, false, -1, a))) return rc;
    bool head_done = false;
    if (e.bf16 && (rc = add_head_tc(e, a, h, c.out_channels, &head_done))) return rc;
    if (!head_done && (rc = add_conv(e, "out.2", "out.2", 3, 1, 0, a, -1, false, h.C, h.H, h.W, c.out_channels, "", -1, -1, 0, -1, -1, -1, -1, true))) return rc;
  }

  // ---- embedding path weights ----
  const float* p = nullptr;
  if ((rc = fetch(e, "time_embed.0.weight", (int64_t)e.ted * mc, &p))) return rc; if ((rc = upload(e, p, (size_t)e.ted * mc, &e.w_t1))) return rc;
  if ((rc = fetch(e, "time_embed.0.bias", e.ted, &p))) return rc;                 if ((rc = upload(e, p, e.ted, &e.b_t1))) return rc;
  if ((rc = fetch(e, "time_embed.2.weight", (int64_t)e.ted * e.ted, &p))) return rc; if ((rc = upload(e, p, (size_t)e.ted * e.ted, &e.w_t2))) return rc;
  if ((rc = fetch(e, "time_embed.2.bias", e.ted, &p))) return rc;                 if ((rc = upload(e, p, e.ted, &e.b_t2))) return rc;
  if (c.num_classes > 0) {
    if ((rc = fetch(e, "label_emb.weight", (int64_t)c.num_classes * e.ted, &p))) return rc;
    if ((rc = upload(e, p, (size_t)c.num_classes * e.ted, &e.label_emb))) return rc;
  }
  std::vector<float> wcat((size_t)e.emb_total * e.ted), bcat(e.emb_total);
  size_t row = 0;
  for (auto& pc : emb_pieces) {
    const float *w = nullptr, *b = nullptr;
    if ((rc = fetch(e, pc.prefix + ".weight", (int64_t)pc.width * e.ted, &w))) return rc;
    if ((rc = fetch(e, pc.prefix + ".bias", pc.width, &b))) return rc;
    std::memcpy(&wcat[row * e.ted], w, sizeof(float) * pc.width * e.ted);
    std::memcpy(&bcat[row], b, sizeof(float) * pc.width);
    row += pc.width;
  }
  if ((rc = upload(e, wcat.data(), wcat.size(), &e.w_emb_cat))) return rc;
  if ((rc = upload(e, bcat.data(), bcat.size(), &e.b_emb_cat))) return rc;
  e.flops_per_sample = 2.0 * ((double)mc * e.ted + (double)e.ted * e.ted + (double)e.emb_total * e.ted);
  for (auto& op : e.ops) e.flops_per_sample += op.flops;

  // ---- fold single-source GroupNorms into the conv that produces their input (a block's last conv, a down-sampling
  // conv): the conv's epilogue then writes the un-normalised block output AND the normalised tensor, and the GroupNorm
  // pass with its extra read of the block output disappears ----
  if (e.bf16 && !(e.cfg.flags & CFM_FLAG_SEPARATE_GROUPNORM)) {
    for (size_t i = 1; i < e.ops.size(); ++i) {
      if (e.ops[i].kind != OP_GN || e.ops[i - 1].kind != OP_CONV) continue;
      if (!tc_conv_attach_gn(e, e.ops[i - 1], e.ops[i])) continue;
      e.ops[i - 1].name += "+" + e.ops[i].name;
      e.ops.erase(e.ops.begin() + i);
      --i;
    }
  }

  // ---- liveness + first-fit arena assignment (per-sample element offsets, 64-element aligned) ----
  for (int i = 0; i < (int)e.ops.size(); ++i) {
    const Op& op = e.ops[i];
    for (int id : {op.src0, op.src1, op.skip0, op.skip1, op.res0, op.res1})
      if (id >= 0) e.tensors[id].last_use = i;
    for (int o : {op.out, op.out2})
      if (o >= 0) { if (e.tensors[o].first_use < 0) e.tensors[o].first_use = i; e.tensors[o].last_use = std::max(e.tensors[o].last_use, i); }
  }
  struct Blk { long long off, len; };
  std::vector<Blk> free_list;
  long long top = 0;
  auto align = [](long long v) { return (v + 63) / 64 * 64; };
  for (int i = 0; i < (int)e.ops.size(); ++i) {
    for (const int o : {e.ops[i].out, e.ops[i].out2})
    if (o >= 0 && e.tensors[o].first_use == i) {
      const long long need = align(e.tensors[o].elems());
      int best = -1;
      for (int k = 0; k < (int)free_list.size(); ++k)
        if (free_list[k].len >= need && (best < 0 || free_list[k].len < free_list[best].len)) best = k;
      if (best >= 0) {
        e.tensors[o].off = free_list[best].off;
        free_list[best].off += need; free_list[best].len -= need;
        if (free_list[best].len == 0) free_list.erase(free_list.begin() + best);
      } else { e.tensors[o].off = top; top += need; }
    }
    for (int t = 0; t < (int)e.tensors.size(); ++t) {
      if (e.tensors[t].last_use == i && e.tensors[t].off >= 0) {
        Blk b{e.tensors[t].off, align(e.tensors[t].elems())};
        free_list.push_back(b);
        std::sort(free_list.begin(), free_list.end(), [](const Blk& a, const Blk& c2) { return a.off < c2.off; });
        for (int k = 0; k + 1 < (int)free_list.size();)   // coalesce
          if (free_list[k].off + free_list[k].len == free_list[k + 1].off) { free_list[k].len += free_list[k + 1].len; free_list.erase(free_list.begin() + k + 1); }
          else ++k;
      }
    }
  }
  e.arena_elems_per_sample = top;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// execution
// ---------------------------------------------------------------------------------------------
static size_t esize(const Engine& e) { return e.bf16 ? 2 : 4; }

static void drop_graphs(Engine& e) {
  for (auto& kv : e.graphs) cudaGraphExecDestroy(kv.second);
  e.graphs.clear();
  e.graph_nodes.clear();
  e.graph_lru.clear();
}

static int ensure_batch(Engine& e, int B) {
  if (B > e.arena_batch) {
    if (e.arena) { cudaFree(e.arena); e.arena = nullptr; }
    if (e.gn_exch) { cudaFree(e.gn_exch); e.gn_exch = nullptr; }
    tc_conv_release(e);   // tensor maps hold arena addresses
    attn_tc_release(e);
    attn_flash_release(e);
    attn_wide_release(e);
    drop_graphs(e);
    const size_t bytes = (size_t)e.arena_elems_per_sample * B * esize(e);
    CU_CHECK(e, cudaMalloc(&e.arena, bytes));
    if (e.gn_tiles_per_sample > 0) {
      CU_CHECK(e, cudaMalloc((void**)&e.gn_exch, sizeof(uint2) * 64 * (size_t)e.gn_tiles_per_sample * B));
      CU_CHECK(e, cudaMemset(e.gn_exch, 0, sizeof(uint2) * 64 * (size_t)e.gn_tiles_per_sample * B));      // epoch 0 = never written
      if (!e.gn_epoch) { CU_CHECK(e, cudaMalloc((void**)&e.gn_epoch, sizeof(unsigned))); CU_CHECK(e, cudaMemset(e.gn_epoch, 0, sizeof(unsigned))); }
    }
    e.arena_batch = B;
  }
  const int rows_needed = std::max(B, std::max(1, e.cfg.num_classes));
  if (rows_needed > e.rows_cap) {
    drop_graphs(e);
    for (void* p : {(void*)e.t_rows, (void*)e.hidden, (void*)e.semb, (void*)e.emb_out, (void*)e.label_idx, (void*)e.row_of_sample})
      if (p) cudaFree(p);
    CU_CHECK(e, cudaMalloc(&e.t_rows, sizeof(float) * rows_needed));
    CU_CHECK(e, cudaMalloc(&e.hidden, sizeof(float) * (size_t)rows_needed * e.ted));
    CU_CHECK(e, cudaMalloc(&e.semb, sizeof(float) * (size_t)rows_needed * e.ted));
    CU_CHECK(e, cudaMalloc(&e.emb_out, sizeof(float) * (size_t)rows_needed * std::max(e.emb_total, 1)));
    CU_CHECK(e, cudaMalloc(&e.label_idx, sizeof(long long) * rows_needed));
    CU_CHECK(e, cudaMalloc(&e.row_of_sample, sizeof(int) * rows_needed));
    e.rows_cap = rows_needed;
  }
  return 0;
}

void* tensor_ptr(const Engine& e, int id, int B) {
  if (id < 0) return nullptr;
  return (char*)e.arena + (size_t)e.tensors[id].off * B * esize(e);
}

// rows of the embedding table: uniform t -> one row (or one per class); per-sample t -> one per sample
__global__ void setup_rows_kernel(int B, int rows, int uniform, float t_scalar, const float* t_dev,
                                  const float* t_table, const int* step_counter,
                                  const long long* y, float* t_rows, long long* label_idx, int* row_of_sample, unsigned* epoch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && epoch) *epoch += 1u;                   // one U-Net evaluation = one epoch of the fused-GroupNorm exchange flags
  if (t_table) t_scalar = t_table[*step_counter];      // sampler loops: time of the current step lives on the device
  if (i < rows) {
    t_rows[i] = uniform ? t_scalar : t_dev[i];
    label_idx[i] = uniform ? (long long)i : (y ? y[i] : 0);
  }
  if (i < B) row_of_sample[i] = uniform ? (y ? (int)y[i] : 0) : i;
}

// fp32 convs whose channel counts are multiples of 16 (all but the stem and the head) take the register-blocked kernel
static bool launch_conv_fast(const ConvArgs<float>& a, long long M, cudaStream_t st) {
  if (a.src_nchw0 || a.out_nchw || !a.out) return false;
  if (a.C0 % 16 || a.C1 % 16 || a.S0 % 16 || a.S1 % 16 || a.Cout % 64 || a.R0 % 4 || a.R1 % 4 || (a.res0 && a.R0 + a.R1 != a.Cout)) return false;
  const unsigned gm = (unsigned)((M + CF_BM - 1) / CF_BM);
  if (a.Cout % 128 == 0 && (long long)gm * (a.Cout / 128) >= 148) conv_fp32_fast_kernel<128><<<dim3(gm, a.Cout / 128), 256, 0, st>>>(a);
  else conv_fp32_fast_kernel<64><<<dim3(gm, a.Cout / 64), 256, 0, st>>>(a);
  return true;
}
static bool launch_conv_fast(const ConvArgs<bf16>&, long long, cudaStream_t) { return false; }

template <typename T>
static int run_ops(Engine& e, int B, const float* x, const float* cond, float* out, cudaStream_t st) {
  int op_index = 0;
  for (const Op& op : e.ops) {
    if (e.profiling) cudaEventRecord(e.prof_events[op_index], st);
    ++op_index;
    switch (op.kind) {
      case OP_CONV: {
        if (op.tc) {
          int rc = tc_conv_launch(e, op, B, st, out);
          if (rc) return rc;
          e.launches++;
          break;
        }
        if (head_conv_supported(e, op)) {
          int rc = head_conv_launch(e, op, B, out, st);
          if (rc) return rc;
          e.launches++;
          break;
        }
        if (stem_conv_supported(e, op)) {
          int rc = stem_conv_launch(e, op, B, x, cond, st);
          if (rc) return rc;
          e.launches++;
          break;
        }
        ConvArgs<T> a{};
        if (op.src_is_input) {
          const int cx = cond ? e.x_channels() : e.cfg.in_channels;
          a.src_nchw0 = x; a.C0 = cx; a.src_nchw1 = cond; a.C1 = e.cfg.in_channels - cx;
        } else {
          a.src0 = (const T*)tensor_ptr(e, op.src0, B); a.C0 = e.tensors[op.src0].C;
          a.src1 = (const T*)tensor_ptr(e, op.src1, B); a.C1 = op.src1 >= 0 ? e.tensors[op.src1].C : 0;
        }
        a.Hin = op.Hin; a.Win = op.Win; a.ups = op.ups; a.stride = op.stride; a.ks = op.ks; a.w_main = op.w_main;
        a.skip0 = (const T*)tensor_ptr(e, op.skip0, B); a.S0 = op.skip0 >= 0 ? e.tensors[op.skip0].C : 0;
        a.skip1 = (const T*)tensor_ptr(e, op.skip1, B); a.S1 = op.skip1 >= 0 ? e.tensors[op.skip1].C : 0;
        a.w_skip = op.w_skip; a.bias = op.bias;
        if (op.emb_off >= 0) { a.emb = e.emb_out + op.emb_off; a.emb_stride = e.emb_total; a.emb_row = e.row_of_sample; }
        a.res0 = (const T*)tensor_ptr(e, op.res0, B); a.R0 = op.res0 >= 0 ? e.tensors[op.res0].C : 0;
        a.res1 = (const T*)tensor_ptr(e, op.res1, B); a.R1 = op.res1 >= 0 ? e.tensors[op.res1].C : 0;
        if (op.out_is_output) a.out_nchw = out; else a.out = (T*)tensor_ptr(e, op.out, B);
        a.B = B; a.Hout = op.Hout; a.Wout = op.Wout; a.Cout = op.Cout;
        const long long M = (long long)B * op.Hout * op.Wout;
        if (launch_conv_fast(a, M, st)) { e.launches++; break; }     // exact mode: register-blocked fp32 kernel
        dim3 grid((unsigned)((M + CG_BM - 1) / CG_BM), (unsigned)((op.Cout + CG_BN - 1) / CG_BN));
        conv_generic_kernel<T><<<grid, 256, 0, st>>>(a);
        e.launches++;
        break;
      }
      case OP_GN: {
        if (e.bf16 && gn_bf16_supported(e, op)) {
          int rc = gn_bf16_launch(e, op, B, st);
          if (rc) return rc;
          e.launches++;
          break;
        }
        GnArgs<T> a{};
        a.src0 = (const T*)tensor_ptr(e, op.src0, B); a.C0 = e.tensors[op.src0].C;
        a.src1 = (const T*)tensor_ptr(e, op.src1, B); a.C1 = op.src1 >= 0 ? e.tensors[op.src1].C : 0;
        a.HW = op.Hin * op.Win; a.gamma = op.gamma; a.beta = op.beta; a.eps = 1e-5f; a.silu = op.silu;
        if (op.film) { a.film = e.emb_out + op.emb_off; a.film_stride = e.emb_total; a.film_row = e.row_of_sample; }
        a.out = (T*)tensor_ptr(e, op.out, B);
        a.exact = e.bf16 ? 0 : 1;
        const int n = (op.Cin / 32) * a.HW;
        constexpr int kGnSmemBytes = 200 * 1024;
        static DeviceOnce gn_attr_set;
        if (gn_attr_set.pending(e.device)) {
          CU_CHECK(e, cudaFuncSetAttribute(groupnorm_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGnSmemBytes));
          CU_CHECK(e, cudaFuncSetAttribute(groupnorm_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGnSmemBytes));
          gn_attr_set.done(e.device);
        }
        const int cap = kGnSmemBytes / 4;
        a.smem_elems = n <= cap ? n : 0;
        groupnorm_kernel<T><<<B * 32, 256, (size_t)a.smem_elems * 4, st>>>(a);
        e.launches++;
        break;
      }
      case OP_HEAD_GATHER: {
        int rc = head_gather_launch(e, op, B, out, st);
        if (rc) return rc;
        e.launches++;
        break;
      }
      case OP_IM2COL: {
        int rc = stem_im2col_launch(e, op, B, x, cond, st);
        if (rc) return rc;
        e.launches++;
        break;
      }
      case OP_RESAMPLE: {
        if (resample_bf16_supported(e, op)) {
          int rc = resample_bf16_launch(e, op, B, st);
          if (rc) return rc;
          e.launches++;
          break;
        }
        const long long total = (long long)B * e.tensors[op.out].elems();
        const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)e.sm_count * 16);
        resample_kernel<T><<<blocks, 256, 0, st>>>((const T*)tensor_ptr(e, op.src0, B), (T*)tensor_ptr(e, op.out, B), B, op.Hin, op.Win, op.Cin, op.up);
        e.launches++;
        break;
      }
      case OP_ATTN: {
        if (attn_tc_supported(e, op)) {
          int rc = attn_tc_launch(e, op, B, st);
          if (rc) return rc;
          e.launches++;
          break;
        }
        if (attn_flash_supported(e, op)) {
          int rc = attn_flash_launch(e, op, B, st);
          if (rc) return rc;
          e.launches++;
          break;
        }
        if (attn_wide_supported(e, op)) {
          int rc = attn_wide_launch(e, op, B, st);
          if (rc) return rc;
          e.launches++;
          break;
        }
        AttnArgs<T> a{(const T*)tensor_ptr(e, op.src0, B), (T*)tensor_ptr(e, op.out, B), B, op.Hin * op.Win, op.heads, op.ch, e.cfg.use_new_attention_order};
        const int nw = 8;
        dim3 grid((a.T_len + nw - 1) / nw, B * op.heads);
        const size_t smem = sizeof(float) * (32 * (op.ch + 1) + 32 * op.ch);
        const int cpl = (op.ch + 31) / 32;
        if (smem > 200 * 1024 || cpl > 16) return fail(e, CFM_ERR_INVALID, "attention head width too large for the generic kernel");
        if (smem > 48 * 1024) {
          cudaFuncSetAttribute(attention_generic_kernel<T, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
          cudaFuncSetAttribute(attention_generic_kernel<T, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        }
        if (cpl <= 1) attention_generic_kernel<T, 1><<<grid, nw * 32, smem, st>>>(a);
        else if (cpl <= 2) attention_generic_kernel<T, 2><<<grid, nw * 32, smem, st>>>(a);
        else if (cpl <= 4) attention_generic_kernel<T, 4><<<grid, nw * 32, smem, st>>>(a);
        else if (cpl <= 8) attention_generic_kernel<T, 8><<<grid, nw * 32, smem, st>>>(a);
        else attention_generic_kernel<T, 16><<<grid, nw * 32, smem, st>>>(a);
        e.launches++;
        break;
      }
    }
  }
  if (e.profiling) cudaEventRecord(e.prof_events[op_index], st);
  return 0;
}

// drop_labels: evaluate a class-conditional model WITHOUT its label embedding (the unconditional branch of
// classifier-free guidance; y must then be NULL).
static int forward_impl(Engine& e, int B, const float* x, const float* cond, const float* t_dev, float t_scalar,
                        const int64_t* y, float* out, cudaStream_t st, const float* t_table = nullptr,
                        const int* step_counter = nullptr, bool drop_labels = false) {
  if (B <= 0) return fail(e, CFM_ERR_INVALID, "batch must be positive");
  if (!x || !out) return fail(e, CFM_ERR_INVALID, "x_dev and out_dev must be non-NULL");
  if (drop_labels ? (y != nullptr) : ((y != nullptr) != (e.cfg.num_classes > 0)))
    return fail(e, CFM_ERR_INVALID, "must specify y if and only if the model is class-conditional");
  if (!cond && e.cfg.in_channels != e.x_channels() && false) return fail(e, CFM_ERR_INVALID, "cond required");
  int rc = ensure_batch(e, B);
  if (rc) return rc;
  const int uniform = t_dev == nullptr;
  const int rows = uniform ? std::max(1, e.cfg.num_classes) : B;
  const int n = std::max(rows, B);
  setup_rows_kernel<<<(n + 255) / 256, 256, 0, st>>>(B, rows, uniform, t_scalar, t_dev, t_table, step_counter, (const long long*)y, e.t_rows, e.label_idx, e.row_of_sample, e.gn_epoch);
  time_hidden_kernel<<<dim3(rows, (e.ted + 7) / 8), 256, sizeof(float) * e.cfg.model_channels, st>>>(e.t_rows, e.cfg.model_channels, e.ted, e.w_t1, e.b_t1, e.hidden);
  linear_rows_kernel<<<dim3((e.ted + 7) / 8, rows), 256, 0, st>>>(e.hidden, e.ted, e.w_t2, e.b_t2, e.ted, drop_labels ? nullptr : e.label_emb, e.label_idx, 1, e.semb);
  linear_rows_kernel<<<dim3((e.emb_total + 7) / 8, rows), 256, 0, st>>>(e.semb, e.ted, e.w_emb_cat, e.b_emb_cat, e.emb_total, nullptr, nullptr, 0, e.emb_out);
  e.launches += 4;
  rc = e.bf16 ? run_ops<bf16>(e, B, x, cond, out, st) : run_ops<float>(e, B, x, cond, out, st);
  if (rc) return rc;
  CU_CHECK(e, cudaGetLastError());
  return 0;
}

static int ew_blocks(const Engine& e, long long n) {
  return (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)e.sm_count * 8));
}

template <typename T>
static int grow(Engine& e, T** p, long long* cap, long long n) {
  if (n <= *cap) return 0;
  drop_graphs(e);
  if (*p) cudaFree(*p);
  *p = nullptr;
  CU_CHECK(e, cudaMalloc((void**)p, sizeof(T) * (size_t)n));
  *cap = n;
  return 0;
}

// Engine-owned sampler buffers: the step graphs only ever reference these, so a captured step can be
// replayed for any caller buffers (state is copied in/out around the loop).
static int ensure_sampler(Engine& e, long long n, long long n_cond, int B, int n_steps) {
  int rc;
  if ((rc = grow(e, &e.v_buf, &e.v_cap, n))) return rc;
  if ((rc = grow(e, &e.x_work, &e.x_cap, n))) return rc;
  if ((rc = grow(e, &e.cond_work, &e.cond_cap, std::max<long long>(n_cond, 1)))) return rc;
  if ((rc = grow(e, &e.img_work, &e.img_cap, n))) return rc;
  if ((rc = grow(e, &e.y_work, &e.y_cap, (long long)B))) return rc;
  if ((rc = grow(e, &e.t_table, &e.t_cap, (long long)std::max(n_steps, 1)))) return rc;
  if ((rc = grow(e, &e.dt_table, &e.dt_cap, (long long)std::max(n_steps, 1)))) return rc;
  if ((rc = grow(e, &e.ddpm_table, &e.ddpm_cap, (long long)std::max(n_steps, 1)))) return rc;
  if (!e.step_counter) CU_CHECK(e, cudaMalloc((void**)&e.step_counter, sizeof(int)));
  if (!e.sampler_params) CU_CHECK(e, cudaMalloc((void**)&e.sampler_params, sizeof(SamplerParams)));
  return 0;
}

// Per-call pointers / seed of the sampler kernels go through device memory, never through captured launch arguments.
static int set_sampler_params(Engine& e, const float* noise, unsigned long long seed, float* traj, cudaStream_t st) {
  const SamplerParams h{noise, seed, traj};
  CU_CHECK(e, cudaMemcpyAsync(e.sampler_params, &h, sizeof(h), cudaMemcpyHostToDevice, st));   // pageable source: staged before return
  return 0;
}

__global__ void counter_add_kernel(int* c) { *c += 1; }

// Runs `body(stream)` n_steps times: directly, or as one captured-and-cached CUDA graph replayed per step.
template <typename Body>
static int run_steps(Engine& e, const std::string& key, bool use_graph, int n_steps, cudaStream_t st, Body body) {
  if (n_steps <= 0) return 0;
  if (!use_graph) {
    for (int k = 0; k < n_steps; ++k) { int rc = body(st); if (rc) return rc; }
    return 0;
  }
  auto it = e.graphs.find(key);
  if (it == e.graphs.end()) {
    cudaStream_t cap = nullptr;
    CU_CHECK(e, cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    cudaError_t ce = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
    if (ce != cudaSuccess) { cudaStreamDestroy(cap); return fail(e, CFM_ERR_CUDA, std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(ce)); }
    const int saved_launches = e.launches;
    int rc = body(cap);
    e.launches = saved_launches;
    cudaGraph_t graph = nullptr;
    ce = cudaStreamEndCapture(cap, &graph);
    cudaStreamDestroy(cap);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return fail(e, CFM_ERR_CUDA, std::string("graph capture failed: ") + cudaGetErrorString(ce));
    cudaGraphExec_t exec = nullptr;
    ce = cudaGraphInstantiate(&exec, graph, 0);
    size_t n_nodes = 0;
    cudaGraphGetNodes(graph, nullptr, &n_nodes);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) return fail(e, CFM_ERR_CUDA, std::string("graph instantiate failed: ") + cudaGetErrorString(ce));
    if ((int)e.graphs.size() >= Engine::kMaxGraphs && !e.graph_lru.empty()) {     // evict the least recently used graph
      const std::string old = e.graph_lru.front();
      e.graph_lru.erase(e.graph_lru.begin());
      auto io = e.graphs.find(old);
      if (io != e.graphs.end()) { cudaGraphExecDestroy(io->second); e.graphs.erase(io); }
      e.graph_nodes.erase(old);
    }
    e.graph_nodes[key] = (int)n_nodes;
    it = e.graphs.emplace(key, exec).first;
  }
  e.graph_lru.erase(std::remove(e.graph_lru.begin(), e.graph_lru.end(), key), e.graph_lru.end());
  e.graph_lru.push_back(key);
  for (int k = 0; k < n_steps; ++k) {
    cudaError_t ce = cudaGraphLaunch(it->second, st);
    if (ce != cudaSuccess) return fail(e, CFM_ERR_CUDA, std::string("graph launch failed: ") + cudaGetErrorString(ce));
  }
  e.launches += e.graph_nodes[key] * n_steps;
  return 0;
}

}  // namespace cfm

// =============================================================================================
// C ABI
// =============================================================================================
using namespace cfm;

struct cfm_engine { Engine impl; };

extern "C" {

int cfm_abi_version(void) { return CFM_ABI_VERSION; }

const char* cfm_last_error(const cfm_engine* e) { return e ? e->impl.err.c_str() : g_create_error.c_str(); }

int cfm_engine_create(const cfm_unet_config* cfg, int32_t n_tensors, const char* const* names,
                      const float* const* host_data, const int64_t* numel, int32_t device, cfm_engine** out) {
  if (!cfg || !out || (n_tensors > 0 && (!names || !host_data || !numel))) { g_create_error = "NULL argument"; return CFM_ERR_INVALID; }
  *out = nullptr;
  std::unique_ptr<cfm_engine> h(new cfm_engine());
  Engine& e = h->impl;
  e.cfg = *cfg; e.device = device; e.bf16 = cfg->precision == CFM_PRECISION_BF16;
  auto bad = [&](int code, const std::string& m) { g_create_error = m; return code; };
  if (cfg->n_levels < 1 || cfg->n_levels > CFM_MAX_LEVELS || cfg->n_attention_ds < 0 || cfg->n_attention_ds > CFM_MAX_LEVELS)
    return bad(CFM_ERR_INVALID, "n_levels / n_attention_ds out of range");
  if (cfg->model_channels <= 0 || cfg->model_channels % 32 || cfg->image_size <= 0 || cfg->in_channels <= 0 || cfg->out_channels <= 0 || cfg->num_res_blocks < 1)
    return bad(CFM_ERR_INVALID, "invalid U-Net configuration (model_channels must be a positive multiple of 32)");
  if (cfg->precision != CFM_PRECISION_FP32 && cfg->precision != CFM_PRECISION_BF16) return bad(CFM_ERR_INVALID, "unknown precision");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return bad(CFM_ERR_CUDA, "no CUDA device available: this engine has no CPU fallback");
  if (device < 0 || device >= ndev) return bad(CFM_ERR_INVALID, "device index out of range");
  if (cudaSetDevice(device) != cudaSuccess) return bad(CFM_ERR_CUDA, "cudaSetDevice failed");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bad(CFM_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10) return bad(CFM_ERR_CUDA, std::string("this library is built for sm_100a (B200); found ") + prop.name);
  e.sm_count = prop.multiProcessorCount;
  for (int i = 0; i < n_tensors; ++i) e.sd[names[i]] = HostTensor{host_data[i], numel[i]};
  int rc = build_plan(e);
  if (rc) { g_create_error = e.err; for (void* p : e.owned) cudaFree(p); return rc; }
  e.sd.clear();
  *out = h.release();
  return CFM_OK;
}

void cfm_engine_destroy(cfm_engine* h) {
  if (!h) return;
  Engine& e = h->impl;
  cudaSetDevice(e.device);
  tc_conv_release(e);
  attn_tc_forget(e);
  attn_flash_forget(e);
  attn_wide_forget(e);
  drop_graphs(e);
  for (void* p : {(void*)e.x_work, (void*)e.cond_work, (void*)e.img_work, (void*)e.y_work, (void*)e.t_table, (void*)e.dt_table,
                  (void*)e.ddpm_table, (void*)e.step_counter, (void*)e.sampler_params})
    if (p) cudaFree(p);
  for (void* p : e.owned) cudaFree(p);
  for (cudaEvent_t ev : e.prof_events) cudaEventDestroy(ev);
  for (void* p : {(void*)e.arena, (void*)e.gn_exch, (void*)e.gn_epoch, (void*)e.t_rows, (void*)e.hidden, (void*)e.semb, (void*)e.emb_out, (void*)e.label_idx, (void*)e.row_of_sample, (void*)e.v_buf, (void*)e.v2_buf})
    if (p) cudaFree(p);
  delete h;
}

int64_t cfm_engine_param_count(const cfm_engine* e) { return e ? e->impl.param_count : 0; }
double cfm_engine_flops_per_sample(const cfm_engine* e) { return e ? e->impl.flops_per_sample : 0; }
int64_t cfm_engine_workspace_bytes(const cfm_engine* e, int32_t batch) {
  return e ? (int64_t)e->impl.arena_elems_per_sample * batch * (e->impl.bf16 ? 2 : 4) : 0;
}
int32_t cfm_engine_kernel_launches(const cfm_engine* e) { return e ? e->impl.launches : 0; }
int32_t cfm_engine_tensor_core_convs(const cfm_engine* e) { return e ? e->impl.n_tc_convs : 0; }
int32_t cfm_engine_cached_graphs(const cfm_engine* e) { return e ? (int32_t)e->impl.graphs.size() : 0; }

int cfm_engine_forward(cfm_engine* h, int32_t batch, const float* x_dev, const float* cond_dev, const float* t_dev,
                       float t_scalar, const int64_t* y_dev, float* out_dev, void* stream) {
  if (!h) return CFM_ERR_INVALID;
  Engine& e = h->impl;
  cudaSetDevice(e.device);
  e.launches = 0;
  return forward_impl(e, batch, x_dev, cond_dev, t_dev, t_scalar, y_dev, out_dev, (cudaStream_t)stream);
}

// Per-op device timing of one NFE (CUDA events around every launch; not for use under graph capture).
int cfm_engine_profile_forward(cfm_engine* h, int32_t batch, const float* x_dev, const float* cond_dev, float t_scalar,
                               const int64_t* y_dev, float* out_dev, int32_t repeats, void* stream) {
  if (!h || repeats < 1) return CFM_ERR_INVALID;
  Engine& e = h->impl;
  cudaSetDevice(e.device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n_ev = e.ops.size() + 1;
  while (e.prof_events.size() < n_ev) { cudaEvent_t ev; CU_CHECK(e, cudaEventCreate(&ev)); e.prof_events.push_back(ev); }
  e.prof_ms.assign(e.ops.size(), 0.0);
  int rc = forward_impl(e, batch, x_dev, cond_dev, nullptr, t_scalar, y_dev, out_dev, st);   // warm (tensor maps, arena)
  if (rc) return rc;
  for (int r = 0; r < repeats; ++r) {
    e.profiling = true;
    rc = forward_impl(e, batch, x_dev, cond_dev, nullptr, t_scalar, y_dev, out_dev, st);
    e.profiling = false;
    if (rc) return rc;
    CU_CHECK(e, cudaStreamSynchronize(st));
    for (size_t i = 0; i < e.ops.size(); ++i) {
      float ms = 0.f;
      CU_CHECK(e, cudaEventElapsedTime(&ms, e.prof_events[i], e.prof_events[i + 1]));
      e.prof_ms[i] += ms / repeats;
    }
  }
  return 0;
}

int32_t cfm_engine_profile_count(const cfm_engine* h) { return h ? (int32_t)h->impl.prof_ms.size() : 0; }

// kind: 0 generic conv, 1 groupnorm, 2 resample, 3 generic attention, 4 tcgen05 conv, 5 tcgen05 attention
int cfm_engine_profile_get(const cfm_engine* h, int32_t i, char* name, int32_t name_cap, int32_t* kind, double* ms,
                           double* flops_per_sample) {
  if (!h || i < 0 || i >= (int32_t)h->impl.prof_ms.size()) return CFM_ERR_INVALID;
  const Op& op = h->impl.ops[i];
  if (name && name_cap > 0) { std::strncpy(name, op.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  if (kind) {
    if (op.kind == OP_CONV) *kind = op.tc ? 4 : 0;
    else if (op.kind == OP_IM2COL || op.kind == OP_HEAD_GATHER) *kind = 2;
    else if (op.kind == OP_ATTN) *kind = (attn_tc_supported(h->impl, op) || attn_flash_supported(h->impl, op) || attn_wide_supported(h->impl, op)) ? 5 : 3;
    else *kind = (int)op.kind;
  }
  if (ms) *ms = h->impl.prof_ms[i];
  if (flops_per_sample) *flops_per_sample = op.flops;
  return 0;
}

int cfm_engine_op_info(const cfm_engine* h, int32_t i, int32_t what, double* value) {
  if (!h || !value || i < 0 || i >= (int32_t)h->impl.ops.size()) return CFM_ERR_INVALID;
  const Engine& e = h->impl;
  const Op& op = e.ops[i];
  const double es = e.bf16 ? 2.0 : 4.0;
  auto tb = [&](int id) { return id >= 0 ? (double)e.tensors[id].elems() * es : 0.0; };
  if (what == 0) { *value = (op.kind == OP_CONV && op.tc) ? tc_conv_executed_flops(op) : op.flops; return 0; }
  if (what == 1) {
    double b = tb(op.src0) + tb(op.src1) + tb(op.skip0) + tb(op.skip1) + tb(op.res0) + tb(op.res1) + tb(op.out) + tb(op.out2);
    const int S = e.cfg.image_size;
    if (op.src_is_input || op.kind == OP_IM2COL) b += 4.0 * e.cfg.in_channels * S * S;           // fp32 NCHW network input
    if (op.out_is_output) b += 4.0 * e.cfg.out_channels * S * S;                                 // fp32 NCHW network output
    *value = b;
    return 0;
  }
  return CFM_ERR_INVALID;
}

static int sample_euler_impl(cfm_engine* h, int32_t batch, float* x_dev, float* cond_dev, const int64_t* y_dev,
                             bool guided, float guidance_w, const float* t_host, const float* dt_host, int32_t n_steps,
                             uint32_t flags, float* traj_dev, uint8_t* img_u8_dev, void* stream) {
  if (!h) return CFM_ERR_INVALID;
  Engine& e = h->impl;
  if (guided && e.cfg.num_classes <= 0) return fail(e, CFM_ERR_INVALID, "classifier-free guidance needs a class-conditional model");
  if (!x_dev || (n_steps > 0 && (!t_host || !dt_host)) || n_steps < 0 || batch <= 0) return fail(e, CFM_ERR_INVALID, "bad argument to cfm_sample_euler");
  if ((y_dev != nullptr) != (e.cfg.num_classes > 0)) return fail(e, CFM_ERR_INVALID, "must specify y if and only if the model is class-conditional");
  cudaSetDevice(e.device);
  cudaStream_t st = (cudaStream_t)stream;
  const int S = e.cfg.image_size;
  const long long n = (long long)batch * e.x_channels() * S * S;
  const long long n_cond = cond_dev ? (long long)batch * (e.cfg.in_channels - e.x_channels()) * S * S : 0;
  int rc = ensure_batch(e, batch); if (rc) return rc;
  if ((rc = ensure_sampler(e, n, n_cond, batch, n_steps))) return rc;
  e.launches = 0;
  // stage state and per-step scalars on the device (stream-ordered; host tables are pageable -> copied before return)
  CU_CHECK(e, cudaMemcpyAsync(e.x_work, x_dev, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  if (cond_dev) CU_CHECK(e, cudaMemcpyAsync(e.cond_work, cond_dev, sizeof(float) * n_cond, cudaMemcpyDeviceToDevice, st));
  if (y_dev) CU_CHECK(e, cudaMemcpyAsync(e.y_work, y_dev, sizeof(long long) * batch, cudaMemcpyDeviceToDevice, st));
  if (n_steps > 0) {
    CU_CHECK(e, cudaMemcpyAsync(e.t_table, t_host, sizeof(float) * n_steps, cudaMemcpyHostToDevice, st));
    CU_CHECK(e, cudaMemcpyAsync(e.dt_table, dt_host, sizeof(float) * n_steps, cudaMemcpyHostToDevice, st));
  }
  CU_CHECK(e, cudaMemsetAsync(e.step_counter, 0, sizeof(int), st));
  if ((rc = set_sampler_params(e, nullptr, 0ull, traj_dev, st))) return rc;
  if (traj_dev) CU_CHECK(e, cudaMemcpyAsync(traj_dev, x_dev, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  if (n_steps == 0 && img_u8_dev) { quantize_u8_kernel<<<ew_blocks(e, n), 256, 0, st>>>(img_u8_dev, x_dev, n); e.launches++; }

  const bool drift = (flags & CFM_EULER_COND_DRIFT) && cond_dev;
  const float* cond_w = cond_dev ? e.cond_work : nullptr;
  const int64_t* y_w = y_dev ? (const int64_t*)e.y_work : nullptr;
  if (guided) { int r2 = grow(e, &e.v2_buf, &e.v2_cap, n); if (r2) return r2; }
  auto body = [&](cudaStream_t s2) -> int {
    int r = forward_impl(e, batch, e.x_work, cond_w, nullptr, 0.f, y_w, e.v_buf, s2, e.t_table, e.step_counter);
    if (r) return r;
    if (guided) {   // second evaluation without the label embedding, then v = v_c + w (v_c - v_u)
      r = forward_impl(e, batch, e.x_work, cond_w, nullptr, 0.f, nullptr, e.v2_buf, s2, e.t_table, e.step_counter, true);
      if (r) return r;
      cfg_combine_kernel<<<ew_blocks(e, n), 256, 0, s2>>>(e.v_buf, e.v2_buf, guidance_w, n);
      e.launches++;
    }
    euler_step_kernel<<<ew_blocks(e, n), 256, 0, s2>>>(e.x_work, e.v_buf, e.dt_table, e.step_counter, n_steps, n,
                                                      drift ? e.cond_work : nullptr, n_cond, e.sampler_params,
                                                      img_u8_dev ? e.img_work : nullptr);
    counter_add_kernel<<<1, 1, 0, s2>>>(e.step_counter);
    e.launches += 2;
    return 0;
  };
  char key[160];
  uint32_t w_bits; std::memcpy(&w_bits, &guidance_w, 4);
  snprintf(key, sizeof(key), "euler:%d:%d:%d:%d:%d:%d:%d:%08x", batch, n_steps, cond_dev != nullptr, y_dev != nullptr, (int)drift,
           img_u8_dev != nullptr, (int)guided, guided ? w_bits : 0u);
  const bool use_graph = (flags & CFM_EULER_USE_GRAPH) != 0;
  if (use_graph && n_steps > 0 && !e.graphs.count(key)) {
    // un-captured dry run first: all lazy host-side setup (tensor maps, kernel attributes) happens outside capture
    if ((rc = forward_impl(e, batch, e.x_work, cond_w, nullptr, 0.f, y_w, e.v_buf, st, e.t_table, e.step_counter))) return rc;
  }
  if ((rc = run_steps(e, key, use_graph, n_steps, st, body))) return rc;
  CU_CHECK(e, cudaMemcpyAsync(x_dev, e.x_work, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  if (drift) CU_CHECK(e, cudaMemcpyAsync(cond_dev, e.cond_work, sizeof(float) * n_cond, cudaMemcpyDeviceToDevice, st));
  if (img_u8_dev && n_steps > 0) CU_CHECK(e, cudaMemcpyAsync(img_u8_dev, e.img_work, (size_t)n, cudaMemcpyDeviceToDevice, st));
  CU_CHECK(e, cudaGetLastError());
  return 0;
}

int cfm_sample_euler(cfm_engine* h, int32_t batch, float* x_dev, float* cond_dev, const int64_t* y_dev,
                     const float* t_host, const float* dt_host, int32_t n_steps, uint32_t flags,
                     float* traj_dev, uint8_t* img_u8_dev, void* stream) {
  return sample_euler_impl(h, batch, x_dev, cond_dev, y_dev, false, 0.f, t_host, dt_host, n_steps, flags, traj_dev, img_u8_dev, stream);
}

int cfm_sample_euler_cfg(cfm_engine* h, int32_t batch, float* x_dev, float* cond_dev, const int64_t* y_dev,
                         float guidance_w, const float* t_host, const float* dt_host, int32_t n_steps, uint32_t flags,
                         float* traj_dev, uint8_t* img_u8_dev, void* stream) {
  return sample_euler_impl(h, batch, x_dev, cond_dev, y_dev, true, guidance_w, t_host, dt_host, n_steps, flags, traj_dev, img_u8_dev, stream);
}

// Scalars of chain step i.  unfused = false: the posterior kernel of step i also applies the mask blend of step i - 1
// (one launch per step); unfused = true: the blend of step i runs on its own before the U-Net call (chains with
// corrector steps, and the stepwise API where a Python-level eps network sits between the launches).
static DdpmStepScalars ddpm_scalars(const cfm_ddpm_tables* tb, const cfm_ddpm_options* opt, int i, bool unfused) {
  const int Ns = tb->Ns, n_corr = (int)opt->n_corrector;
  const bool repl = opt->mode == CFM_DDPM_REPLACEMENT;
  auto blend_at = [&](int j) { return repl && j >= 0 && j < opt->replace_below_step; };
  DdpmStepScalars s{};
  s.a = tb->sqrt_recip_alphas_cumprod[i]; s.b = tb->sqrt_recipm1_alphas_cumprod[i];
  s.c1 = tb->posterior_mean_coef1[i]; s.c2 = tb->posterior_mean_coef2[i];
  s.sigma = expf(0.5f * tb->posterior_log_variance_clipped[i]);
  s.add_noise = i > 0;
  s.blend_next = blend_at(i - 1);
  s.noise_condition = opt->noise_condition; s.pad_value = opt->pad_value;
  if (s.blend_next) { s.sa = tb->sqrt_alphas_cumprod[i - 1]; s.sb = tb->sqrt_one_minus_alphas_cumprod[i - 1]; }
  s.final_clip = i == 0 && n_corr == 0;      // with correctors the last corrector of step 0 clips
  s.chain_index = i;
  s.n_slots = 2 + n_corr;
  if (unfused) {
    // un-fused order: [blend i] -> U-Net -> posterior draw -> n_corr x (U-Net -> Langevin step)
    s.blend_cur = blend_at(i); s.blend_next = 0;
    s.sa_cur = tb->sqrt_alphas_cumprod[i]; s.sb_cur = tb->sqrt_one_minus_alphas_cumprod[i];
    s.corr_r = 1.0f / tb->sqrt_one_minus_alphas_cumprod[i];
    const double dt = (1.0 - 0.00001) / Ns;            // (tmax - tmin) / Ns, sde_diffusion.py:130-132
    s.corr_cd = (float)(0.5 * dt * (double)opt->corrector_delta);
    s.corr_cn = (float)std::sqrt(dt * (double)opt->corrector_delta);
  }
  return s;
}

int cfm_sample_ddpm(cfm_engine* h, int32_t batch, float* x_dev, const float* condition_dev,
                    const cfm_ddpm_tables* tb, const cfm_ddpm_options* opt, const float* noise_dev,
                    uint64_t seed, void* stream) {
  if (!h) return CFM_ERR_INVALID;
  Engine& e = h->impl;
  if (!x_dev || !tb || !opt || batch <= 0 || tb->Ns <= 0) return fail(e, CFM_ERR_INVALID, "bad argument to cfm_sample_ddpm");
  if (opt->mode != CFM_DDPM_PRIOR && !condition_dev) return fail(e, CFM_ERR_INVALID, "condition_dev required for conditional sampling");
  if (e.cfg.num_classes > 0) return fail(e, CFM_ERR_INVALID, "class-conditional DDPM sampling is not part of the reference path");
  cudaSetDevice(e.device);
  cudaStream_t st = (cudaStream_t)stream;
  const int S = e.cfg.image_size, Ns = tb->Ns;
  const long long n = (long long)batch * e.x_channels() * S * S;
  const bool repl = opt->mode == CFM_DDPM_REPLACEMENT;
  const bool amort = opt->mode == CFM_DDPM_AMORTIZED;
  const int n_corr = (int)opt->n_corrector;
  if (n_corr < 0 || n_corr > 16) return fail(e, CFM_ERR_INVALID, "n_corrector out of range");
  if (n_corr > 0 && !repl && !amort) return fail(e, CFM_ERR_INVALID, "corrector steps need Replacement or Amortized conditioning");
  const int n_slots = 2 + n_corr;
  int rc = ensure_batch(e, batch); if (rc) return rc;
  if ((rc = ensure_sampler(e, n, condition_dev ? n : 0, batch, Ns))) return rc;
  // Amortized correctors call x0_model without a condition, i.e. with likelihood.none_like(xi): a constant image
  // (sampling.py:34-37, 116); pad_value carries that constant.
  if (amort && n_corr > 0 && (rc = grow(e, &e.v2_buf, &e.v2_cap, n))) return rc;
  const float* corr_cond = amort && n_corr > 0 ? e.v2_buf : nullptr;
  e.launches = 0;
  auto blend_at = [&](int i) { return repl && i >= 0 && i < opt->replace_below_step; };

  // per-step scalars in execution order k = 0..Ns-1  <->  chain index i = Ns-1-k
  std::vector<DdpmStepScalars> tab(Ns);
  std::vector<float> tt(Ns);
  for (int k = 0; k < Ns; ++k) {
    const int i = Ns - 1 - k;
    tab[k] = ddpm_scalars(tb, opt, i, n_corr > 0);
    tt[k] = tb->model_time[i];
  }
  CU_CHECK(e, cudaMemcpyAsync(e.x_work, x_dev, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  if (condition_dev) CU_CHECK(e, cudaMemcpyAsync(e.cond_work, condition_dev, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  CU_CHECK(e, cudaMemcpyAsync(e.ddpm_table, tab.data(), sizeof(DdpmStepScalars) * Ns, cudaMemcpyHostToDevice, st));
  CU_CHECK(e, cudaMemcpyAsync(e.t_table, tt.data(), sizeof(float) * Ns, cudaMemcpyHostToDevice, st));
  CU_CHECK(e, cudaMemsetAsync(e.step_counter, 0, sizeof(int), st));
  if ((rc = set_sampler_params(e, noise_dev, seed, nullptr, st))) return rc;
  if (corr_cond) { fill_f32_kernel<<<ew_blocks(e, n), 256, 0, st>>>(e.v2_buf, opt->pad_value, n); e.launches++; }
  CU_CHECK(e, cudaStreamSynchronize(st));   // host staging vectors go out of scope below
  if (n_corr == 0 && blend_at(Ns - 1)) {
    ddpm_blend_kernel<<<ew_blocks(e, n), 256, 0, st>>>(e.x_work, e.cond_work, tb->sqrt_alphas_cumprod[Ns - 1],
        tb->sqrt_one_minus_alphas_cumprod[Ns - 1], opt->pad_value, opt->noise_condition,
        noise_dev ? noise_dev + ((long long)(Ns - 1) * 2) * n : nullptr, seed, 2u * (Ns - 1), n);
    e.launches++;
  }
  auto body = [&](cudaStream_t s2) -> int {
    if (n_corr > 0 && repl) { ddpm_blend_table_kernel<<<ew_blocks(e, n), 256, 0, s2>>>(e.x_work, e.cond_work, e.ddpm_table, e.step_counter, e.sampler_params, n); e.launches++; }
    int r = forward_impl(e, batch, e.x_work, amort ? e.cond_work : nullptr, nullptr, 0.f, nullptr, e.v_buf, s2, e.t_table, e.step_counter);
    if (r) return r;
    ddpm_step_kernel<<<ew_blocks(e, n), 256, 0, s2>>>(e.x_work, e.v_buf, e.ddpm_table, e.step_counter,
                                                     condition_dev ? e.cond_work : nullptr, e.sampler_params, n);
    for (int c = 0; c < n_corr; ++c) {
      r = forward_impl(e, batch, e.x_work, corr_cond, nullptr, 0.f, nullptr, e.v_buf, s2, e.t_table, e.step_counter);
      if (r) return r;
      ddpm_corrector_kernel<<<ew_blocks(e, n), 256, 0, s2>>>(e.x_work, e.v_buf, e.ddpm_table, e.step_counter, e.sampler_params, 2 + c, c == n_corr - 1, n);
      e.launches++;
    }
    counter_add_kernel<<<1, 1, 0, s2>>>(e.step_counter);
    e.launches += 2;
    return 0;
  };
  char key[160];
  uint32_t d_bits; std::memcpy(&d_bits, &opt->corrector_delta, 4);
  snprintf(key, sizeof(key), "ddpm:%d:%d:%d:%d:%08x", batch, opt->mode, condition_dev != nullptr, n_corr, d_bits);
  if (opt->use_graph && !e.graphs.count(key)) {
    if ((rc = forward_impl(e, batch, e.x_work, amort ? e.cond_work : nullptr, nullptr, 0.f, nullptr, e.v_buf, st, e.t_table, e.step_counter))) return rc;
  }
  if ((rc = run_steps(e, key, opt->use_graph != 0, Ns, st, body))) return rc;
  CU_CHECK(e, cudaMemcpyAsync(x_dev, e.x_work, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  CU_CHECK(e, cudaGetLastError());
  return 0;
}

static int rk_pack(RkPtrs* p, const float* const* k_dev, const float* coef_host, int n_k) {
  if (!k_dev || !coef_host || n_k < 1 || n_k > 8) return CFM_ERR_INVALID;
  p->n_k = n_k;
  for (int j = 0; j < 8; ++j) { p->k[j] = j < n_k ? k_dev[j] : nullptr; p->coef[j] = j < n_k ? coef_host[j] : 0.f; }
  return 0;
}

int cfm_rk_combine(float* out_dev, const float* y_dev, const float* const* k_dev, const float* coef_host,
                   int32_t n_k, float dt, int64_t n, void* stream) {
  RkPtrs p; if (rk_pack(&p, k_dev, coef_host, n_k) || !out_dev || !y_dev || n < 0) return CFM_ERR_INVALID;
  if (n == 0) return 0;
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  rk_combine_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out_dev, y_dev, p, dt, n);
  return cudaGetLastError() == cudaSuccess ? 0 : CFM_ERR_CUDA;
}

// scratch of the deterministic error norm: per device, block partials + the ticket counter (zeroed once; the kernel
// leaves the ticket at zero)
static int rk_scratch(int blocks, double** partial, unsigned** ticket) {
  static std::mutex mu;
  static std::map<int, std::pair<double*, unsigned*>> per_dev;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return CFM_ERR_CUDA;
  std::lock_guard<std::mutex> lk(mu);
  auto it = per_dev.find(dev);
  if (it == per_dev.end()) {
    double* p = nullptr; unsigned* t = nullptr;
    if (cudaMalloc((void**)&p, sizeof(double) * 148 * 8) != cudaSuccess || cudaMalloc((void**)&t, sizeof(unsigned)) != cudaSuccess) return CFM_ERR_OOM;
    if (cudaMemset(t, 0, sizeof(unsigned)) != cudaSuccess) return CFM_ERR_CUDA;
    it = per_dev.emplace(dev, std::make_pair(p, t)).first;
  }
  if (blocks > 148 * 8) return CFM_ERR_INTERNAL;
  *partial = it->second.first; *ticket = it->second.second;
  return 0;
}

int cfm_rk_error_sumsq(double* sumsq_dev, const float* y0_dev, const float* y1_dev, const float* const* k_dev,
                       const float* coef_host, int32_t n_k, float dt, float rtol, float atol, int64_t n, void* stream) {
  RkPtrs p; if (rk_pack(&p, k_dev, coef_host, n_k) || !sumsq_dev || !y0_dev || !y1_dev || n < 0) return CFM_ERR_INVALID;
  if (cudaMemsetAsync(sumsq_dev, 0, sizeof(double), (cudaStream_t)stream) != cudaSuccess) return CFM_ERR_CUDA;
  if (n == 0) return 0;
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  double* partial = nullptr; unsigned* ticket = nullptr;
  if (int rc = rk_scratch(blocks, &partial, &ticket)) return rc;
  rk_error_sumsq_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(sumsq_dev, partial, ticket, y0_dev, y1_dev, p, dt, rtol, atol, n);
  return cudaGetLastError() == cudaSuccess ? 0 : CFM_ERR_CUDA;
}

int cfm_rk_scaled_sumsq(double* sumsq_dev, const float* a_dev, const float* b_dev, const float* y_dev, float rtol, float atol,
                        int64_t n, void* stream) {
  if (!sumsq_dev || !a_dev || !y_dev || n < 0) return CFM_ERR_INVALID;
  if (cudaMemsetAsync(sumsq_dev, 0, sizeof(double), (cudaStream_t)stream) != cudaSuccess) return CFM_ERR_CUDA;
  if (n == 0) return 0;
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  double* partial = nullptr; unsigned* ticket = nullptr;
  if (int rc = rk_scratch(blocks, &partial, &ticket)) return rc;
  rk_scaled_sumsq_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(sumsq_dev, partial, ticket, a_dev, b_dev, y_dev, rtol, atol, n);
  return cudaGetLastError() == cudaSuccess ? 0 : CFM_ERR_CUDA;
}

int cfm_rk_dense_output(float* out_dev, const float* y0_dev, const float* y1_dev, const float* ymid_dev, const float* f0_dev,
                        const float* f1_dev, float dt, float x, int64_t n, void* stream) {
  if (!out_dev || !y0_dev || !y1_dev || !ymid_dev || !f0_dev || !f1_dev || n < 0) return CFM_ERR_INVALID;
  if (n == 0) return 0;
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  rk_dense_output_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out_dev, y0_dev, y1_dev, ymid_dev, f0_dev, f1_dev, dt, x, n);
  return cudaGetLastError() == cudaSuccess ? 0 : CFM_ERR_CUDA;
}

// One launch of a DDPM reverse-chain step for a caller that evaluates the eps network itself (a plain Python callable:
// AD/experiments/main.py:140 passes `lambda xi, i: ema_network(xi, 1.0 * i / ddpm.Ns)`).
//   phase 0: mask blend of step i (Replacement; before the network call)   phase 1: posterior draw from eps
//   phase 2 + c: Langevin corrector c from eps (the network re-evaluated on the current x)
int cfm_ddpm_step(float* x_dev, const float* eps_dev, const float* condition_dev, const cfm_ddpm_tables* tb,
                  const cfm_ddpm_options* opt, int32_t chain_index, int32_t phase, const float* noise_dev, uint64_t seed,
                  int64_t n, void* stream) {
  if (!x_dev || !tb || !opt || tb->Ns <= 0 || chain_index < 0 || chain_index >= tb->Ns || n < 0 || phase < 0 ||
      phase > 1 + (int)opt->n_corrector || (int)opt->n_corrector > 16) return CFM_ERR_INVALID;
  if (phase >= 1 && !eps_dev) return CFM_ERR_INVALID;
  if (opt->mode == CFM_DDPM_REPLACEMENT && !condition_dev) return CFM_ERR_INVALID;
  if (n == 0) return 0;
  const DdpmStepScalars s = ddpm_scalars(tb, opt, chain_index, true);
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (phase == 0) { if (s.blend_cur) ddpm_blend_value_kernel<<<blocks, 256, 0, st>>>(x_dev, condition_dev, s, noise_dev, seed, n); }
  else if (phase == 1) ddpm_step_value_kernel<<<blocks, 256, 0, st>>>(x_dev, eps_dev, s, condition_dev, noise_dev, seed, n);
  else ddpm_corrector_value_kernel<<<blocks, 256, 0, st>>>(x_dev, eps_dev, s, noise_dev, seed, phase, phase == 1 + (int)opt->n_corrector, n);
  return cudaGetLastError() == cudaSuccess ? 0 : CFM_ERR_CUDA;
}

int cfm_sde_em_step(float* x_dev, const float* drift_dev, const float* score_dev, float dt, float sigma,
                    const float* noise_dev, uint64_t seed, uint32_t stream_id, int64_t n, void* stream) {
  if (!x_dev || !drift_dev || n < 0 || !(dt > 0.f)) return CFM_ERR_INVALID;
  if (n == 0) return 0;
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  sde_em_step_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x_dev, drift_dev, score_dev, dt, sigma, sqrtf(dt), noise_dev, seed, stream_id, n);
  return cudaGetLastError() == cudaSuccess ? 0 : CFM_ERR_CUDA;
}

int cfm_ddpm_em_step(float* x_dev, const float* eps_dev, float beta_t, float sigma_t, double dt, const float* noise_dev,
                     uint64_t seed, uint32_t stream_id, int64_t n, void* stream) {
  if (!x_dev || !eps_dev || n < 0 || !(dt > 0.0) || !(sigma_t > 0.f) || beta_t < 0.f) return CFM_ERR_INVALID;
  if (n == 0) return 0;
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  // dt = 1 / Ns and np.sqrt(dt) are Python doubles in the reference, rounded to fp32 when they meet a tensor
  ddpm_em_step_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x_dev, eps_dev, beta_t, sigma_t, sqrtf(beta_t), (float)dt, (float)std::sqrt(dt), noise_dev, seed, stream_id, n);
  return cudaGetLastError() == cudaSuccess ? 0 : CFM_ERR_CUDA;
}

// Euler-Maruyama sampling of dx = (drift(t, x, y) + score(t, x, y)) dt + sigma dW over the fixed grid t_host / dt_host
// (conditional_mnist.ipynb cells 11-12: torchsde.sdeint(SDE(model, score_model), ..., dt=0.01)).  Both engines evaluate
// the same state; score may be NULL.  noise_dev: [n_steps, B*C*H*W] injected normals, else Philox(seed), stream = step.
int cfm_sample_sde(cfm_engine* drift, cfm_engine* score, int32_t batch, float* x_dev, const int64_t* y_dev,
                   const float* t_host, const float* dt_host, int32_t n_steps, float sigma, const float* noise_dev,
                   uint64_t seed, void* stream) {
  if (!drift) return CFM_ERR_INVALID;
  Engine& e = drift->impl;
  if (!x_dev || batch <= 0 || n_steps < 0 || (n_steps > 0 && (!t_host || !dt_host))) return fail(e, CFM_ERR_INVALID, "bad argument to cfm_sample_sde");
  if (score && (score->impl.device != e.device || score->impl.cfg.image_size != e.cfg.image_size || score->impl.cfg.in_channels != e.cfg.in_channels ||
                score->impl.cfg.out_channels != e.cfg.out_channels || (score->impl.cfg.num_classes > 0) != (e.cfg.num_classes > 0)))
    return fail(e, CFM_ERR_INVALID, "drift and score networks must share device, image shape and class conditioning");
  if ((y_dev != nullptr) != (e.cfg.num_classes > 0)) return fail(e, CFM_ERR_INVALID, "must specify y if and only if the model is class-conditional");
  if (e.cfg.in_channels != e.x_channels()) return fail(e, CFM_ERR_INVALID, "cfm_sample_sde takes unconditional or class-conditional networks");
  cudaSetDevice(e.device);
  cudaStream_t st = (cudaStream_t)stream;
  const int S = e.cfg.image_size;
  const long long n = (long long)batch * e.x_channels() * S * S;
  int rc = ensure_batch(e, batch); if (rc) return rc;
  if ((rc = grow(e, &e.v_buf, &e.v_cap, n))) return rc;
  if (score) {
    if ((rc = ensure_batch(score->impl, batch))) { e.err = score->impl.err; return rc; }
    if ((rc = grow(score->impl, &score->impl.v_buf, &score->impl.v_cap, n))) { e.err = score->impl.err; return rc; }
  }
  e.launches = 0;
  for (int k = 0; k < n_steps; ++k) {
    if ((rc = forward_impl(e, batch, x_dev, nullptr, nullptr, t_host[k], y_dev, e.v_buf, st))) return rc;
    if (score) {
      score->impl.launches = 0;
      if ((rc = forward_impl(score->impl, batch, x_dev, nullptr, nullptr, t_host[k], y_dev, score->impl.v_buf, st))) { e.err = score->impl.err; return rc; }
      e.launches += score->impl.launches;
    }
    sde_em_step_kernel<<<ew_blocks(e, n), 256, 0, st>>>(x_dev, e.v_buf, score ? score->impl.v_buf : nullptr, dt_host[k], sigma, sqrtf(dt_host[k]),
                                                       noise_dev ? noise_dev + (long long)k * n : nullptr, seed, (unsigned)k, n);
    e.launches++;
  }
  CU_CHECK(e, cudaGetLastError());
  return 0;
}

int cfm_resize_bilinear(float* out_dev, const float* in_dev, int64_t planes, int32_t h_in, int32_t w_in, int32_t h_out,
                        int32_t w_out, void* stream) {
  if (!out_dev || !in_dev || planes < 0 || h_in <= 0 || w_in <= 0 || h_out <= 0 || w_out <= 0) return CFM_ERR_INVALID;
  const long long n = (long long)planes * h_out * w_out;
  if (n == 0) return 0;
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  resize_bilinear_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out_dev, in_dev, planes, h_in, w_in, h_out, w_out,
                                                                  (float)h_in / (float)h_out, (float)w_in / (float)w_out);
  return cudaGetLastError() == cudaSuccess ? 0 : CFM_ERR_CUDA;
}

int cfm_fid_accumulate(double* sum_dev, double* outer_dev, const float* feats_dev, int64_t n, int32_t dim, void* stream) {
  if (!sum_dev || !outer_dev || !feats_dev || n < 0 || dim <= 0) return CFM_ERR_INVALID;
  if (n == 0) return 0;
  const unsigned t = (unsigned)((dim + 31) / 32);
  fid_accumulate_kernel<<<dim3(t, t), 256, 0, (cudaStream_t)stream>>>(sum_dev, outer_dev, feats_dev, n, dim);
  return cudaGetLastError() == cudaSuccess ? 0 : CFM_ERR_CUDA;
}

int cfm_make_box_condition(float* cond_dev, const float* images_dev, const int32_t* boxes_dev, int32_t batch,
                           int32_t channels, int32_t height, int32_t width, int32_t patch, float pad_value,
                           int32_t mode, void* stream) {
  if (!cond_dev || !images_dev || !boxes_dev || batch < 0 || channels <= 0 || height <= 0 || width <= 0 || patch < 0 || (mode != 0 && mode != 1)) return CFM_ERR_INVALID;
  const long long n = (long long)batch * channels * height * width;
  if (n == 0) return 0;
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  box_condition_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(cond_dev, images_dev, boxes_dev, batch, channels, height, width, patch, pad_value, mode);
  return cudaGetLastError() == cudaSuccess ? 0 : CFM_ERR_CUDA;
}

int cfm_quantize_u8(uint8_t* out_dev, const float* x_dev, int64_t n, void* stream) {
  if (!out_dev || !x_dev || n < 0) return CFM_ERR_INVALID;
  if (n == 0) return 0;
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  quantize_u8_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out_dev, x_dev, n);
  return cudaGetLastError() == cudaSuccess ? 0 : CFM_ERR_CUDA;
}

}  // extern "C"
