// Internal engine structures (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>
#include "../../include/cfm_b200.h"

namespace cfm {

// A/B switches of the kernels (profiles/README.md) exist only in tuning builds (-DCFM_TUNING, `CFM_BUILD_TUNING=1` for
// __graft_entry__.build): the shipped library reads no environment variable and runs one code path.
inline const char* tuning_env(const char* name) {
#ifdef CFM_TUNING
  return getenv(name);
#else
  (void)name;
  return nullptr;
#endif
}

struct HostTensor { const float* data; int64_t numel; };

// An activation tensor of the plan: NHWC [B, H, W, C]; `off` is a per-sample element offset
// into the arena (multiplied by the batch size at run time).
struct TensorDesc {
  int C = 0, H = 0, W = 0;
  long long off = -1;
  int first_use = -1, last_use = -1;
  long long elems() const { return (long long)C * H * W; }
};

enum OpKind { OP_CONV = 0, OP_GN = 1, OP_RESAMPLE = 2, OP_ATTN = 3, OP_IM2COL = 4, OP_HEAD_GATHER = 5 };

struct TcConvPlan;   // tcgen05 implicit-GEMM lowering of a conv op (conv_tc.cu)

struct Op {
  OpKind kind;
  std::string name;
  // tensor ids (-1 = none). For convs: src = main operand, skip = 1x1 operand, res = identity residual.
  int src0 = -1, src1 = -1, skip0 = -1, skip1 = -1, res0 = -1, res1 = -1, out = -1;
  int out2 = -1;           // conv with a folded GroupNorm that still writes its un-normalised result: out2 = the normalised tensor
  bool src_is_input = false, out_is_output = false;
  bool out_f32 = false;    // tcgen05 conv writing fp32 NHWC rows (the head's per-tap partial products)
  // conv
  int ks = 3, stride = 1, ups = 0, Cin = 0, Cskip = 0, Cout = 0, Hin = 0, Win = 0, Hout = 0, Wout = 0;
  float* w_main = nullptr; float* w_skip = nullptr; float* bias = nullptr;   // fp32 device
  int emb_off = -1;        // offset of this block's vector in the embedding table (conv add / FiLM)
  // groupnorm
  float* gamma = nullptr; float* beta = nullptr; int silu = 0; bool film = false;
  // resample
  int up = 0;
  // attention
  int heads = 0, ch = 0;
  // tensor-core lowering (bf16 mode), null when the generic kernel runs this op
  TcConvPlan* tc = nullptr;
  double flops = 0;        // 2*MAC per sample
  // conv with the following GroupNorm (+SiLU) applied in its epilogue (ResBlock conv1 + out_layers.0; conv_tc.cu, kGN):
  // gn_request is set by the plan (gamma / beta / silu then describe that GroupNorm), gn_fused by tc_conv_prepare.
  bool gn_request = false, gn_fused = false;
  int gn_ctas = 1;              // CTA tiles per sample; > 1: partial sums are exchanged through Engine::gn_exch
  long long gn_exch_off = -1;   // offset of this op's exchange slots, in tiles per sample
};

struct Engine {
  cfm_unet_config cfg{};
  int device = 0;
  bool bf16 = false;
  std::string err;
  std::map<std::string, HostTensor> sd;
  std::vector<void*> owned;            // device allocations freed at destroy
  std::vector<TensorDesc> tensors;
  std::vector<Op> ops;
  long long arena_elems_per_sample = 0;
  void* arena = nullptr; int arena_batch = 0;
  // fused GroupNorm epilogues whose samples span several CTA tiles: per-(sample, tile) partial sums and epoch flags
  uint2* gn_exch = nullptr; unsigned* gn_epoch = nullptr; long long gn_tiles_per_sample = 0;
  // embedding path
  int ted = 0, emb_total = 0;
  float *w_t1 = nullptr, *b_t1 = nullptr, *w_t2 = nullptr, *b_t2 = nullptr, *label_emb = nullptr;
  float *w_emb_cat = nullptr, *b_emb_cat = nullptr;
  float *t_rows = nullptr, *hidden = nullptr, *semb = nullptr, *emb_out = nullptr;
  long long* label_idx = nullptr; int* row_of_sample = nullptr; int rows_cap = 0;
  // scratch for samplers
  float* v_buf = nullptr; long long v_cap = 0;
  float* v2_buf = nullptr; long long v2_cap = 0;      // unconditional velocity of classifier-free guidance
  float* x_work = nullptr; long long x_cap = 0;
  float* cond_work = nullptr; long long cond_cap = 0;
  uint8_t* img_work = nullptr; long long img_cap = 0;
  long long* y_work = nullptr; long long y_cap = 0;
  float* t_table = nullptr; long long t_cap = 0;
  float* dt_table = nullptr; long long dt_cap = 0;
  struct DdpmStepScalars* ddpm_table = nullptr; long long ddpm_cap = 0;
  int* step_counter = nullptr;
  struct SamplerParams* sampler_params = nullptr;     // device: noise pointer, seed, trajectory buffer of the running loop
  std::map<std::string, cudaGraphExec_t> graphs;      // one captured step per sampler configuration (LRU, <= kMaxGraphs)
  std::map<std::string, int> graph_nodes;
  std::vector<std::string> graph_lru;                 // least recently used first
  static constexpr int kMaxGraphs = 8;
  int64_t param_count = 0;
  double flops_per_sample = 0;
  int launches = 0;
  int n_tc_convs = 0;
  int sm_count = 148;
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;
  std::vector<double> prof_ms;
  int x_channels() const { return cfg.out_channels; }
};

// Launch configuration with an optional cluster dimension and programmatic dependent launch (PDL).  A kernel launched
// with pdl = true may start while its predecessor in the stream is still draining; it must execute pdl_wait()
// (griddepcontrol.wait) before its first access to global memory that a predecessor writes or reads.
struct LaunchCfg {
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[2];
  LaunchCfg(dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster, bool pdl) {
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    int n = 0;
    if (cluster > 1) {
      attr[n].id = cudaLaunchAttributeClusterDimension;
      attr[n].val.clusterDim.x = cluster; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
      ++n;
    }
    if (pdl) {
      attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[n].val.programmaticStreamSerializationAllowed = 1;
      ++n;
    }
    cfg.attrs = attr; cfg.numAttrs = n;
  }
};
inline bool pdl_enabled() {
  static const bool on = [] { const char* v = tuning_env("CFM_DISABLE_PDL"); return !(v && v[0] == '1'); }();
  return on;
}

// conv_tc.cu
bool tc_conv_supported(const Engine& e, const Op& op);
int  tc_conv_prepare(Engine& e, Op& op, const std::vector<float>& w_main_oihw, const std::vector<float>& w_skip_oi, int force_block_n = 0);
int  tc_conv_launch(Engine& e, const Op& op, int B, cudaStream_t st, float* out_nchw = nullptr);
bool tc_conv_attach_gn(Engine& e, Op& conv, const Op& gn);
void tc_conv_release(Engine& e);
double tc_conv_executed_flops(const Op& op);   // 2*MAC per sample the tcgen05 kernel issues (padding and folding included)
// bf16 fast kernels (kernels_bf16.cu)
int  gn_bf16_launch(Engine& e, const Op& op, int B, cudaStream_t st);
// attn_tc.cu
bool attn_tc_supported(const Engine& e, const Op& op);
int  attn_tc_launch(Engine& e, const Op& op, int B, cudaStream_t st);
void attn_tc_release(Engine& e);
void attn_tc_forget(Engine& e);
// attn_flash.cu
bool attn_flash_supported(const Engine& e, const Op& op);
int  attn_flash_launch(Engine& e, const Op& op, int B, cudaStream_t st);
void attn_flash_release(Engine& e);
void attn_flash_forget(Engine& e);
// attn_wide.cu
bool attn_wide_supported(const Engine& e, const Op& op);
int  attn_wide_launch(Engine& e, const Op& op, int B, cudaStream_t st);
void attn_wide_release(Engine& e);
void attn_wide_forget(Engine& e);
// cudaFuncSetAttribute acts on the current device's context: remember what was set per device, not per process
// (one process may own engines on several GPUs).
struct DeviceOnce {
  unsigned long long mask = 0;
  bool pending(int dev) const { return dev < 0 || dev >= 64 || !((mask >> dev) & 1ull); }
  void done(int dev) { if (dev >= 0 && dev < 64) mask |= 1ull << dev; }
};

bool gn_bf16_supported(const Engine& e, const Op& op);
bool resample_bf16_supported(const Engine& e, const Op& op);
int  resample_bf16_launch(Engine& e, const Op& op, int B, cudaStream_t st);
bool head_conv_supported(const Engine& e, const Op& op);
int  head_conv_launch(Engine& e, const Op& op, int B, float* out, cudaStream_t st);
bool stem_conv_supported(const Engine& e, const Op& op);
int  stem_conv_launch(Engine& e, const Op& op, int B, const float* x, const float* cond, cudaStream_t st);
int  stem_im2col_launch(Engine& e, const Op& op, int B, const float* x, const float* cond, cudaStream_t st);
int  head_gather_launch(Engine& e, const Op& op, int B, float* out, cudaStream_t st);

}  // namespace cfm
