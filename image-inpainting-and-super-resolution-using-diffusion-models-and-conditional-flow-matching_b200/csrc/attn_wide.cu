// Attention core for WIDE heads on tcgen05/TMEM: head width D = 128 .. 512 (multiples of 64), up to 256 tokens.
// The super-resolution U-Net (num_heads = 1) attends over 16x16 maps with one 384-wide head and over 8x8 maps
// with one 512-wide head (unet.py:424-483) - shapes outside attn_tc.cu (D = 64) and attn_flash.cu (D <= 64).
//
// One CTA per (sample, head, 128-query tile), warp-specialised:
//   warp 0   TMA producer: streams 64-channel chunks through a 3-stage ring - first (Q_c, K_c) pairs, then V pieces
//   warp 1   MMA issuer:   S = sum_c Q_c K_c^T  (128 x 256 x 64 per chunk)          -> TMEM columns [0, 256)
//                          O[:, 64 p .. +64) = P V_p  (V piece as an MN-major operand) -> TMEM columns [0, D), which
//                          re-use the S columns once every softmax thread has read its row
//   warps 2-5  softmax (exact, two passes over the row in TMEM, fp32) -> bf16 P in swizzled shared memory,
//              then the epilogue: O / rowsum -> bf16, 256-bit stores
// TMEM: 512 columns (D = 512 needs them all), so one CTA per SM; the ring keeps the tensor pipe fed.
#include <cstring>
#include <map>
#include <mutex>
#include "engine.h"
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cfm {

void* tensor_ptr(const Engine& e, int id, int B);

constexpr int AW_M = 128, AW_T = 256, AW_STAGES = 3, AW_STAGE_BYTES = 49152;
constexpr int AW_P_OFF = AW_STAGES * AW_STAGE_BYTES;          // 144 KB
constexpr int AW_BAR_OFF = AW_P_OFF + 65536;                  // 208 KB
constexpr int AW_SMEM = AW_BAR_OFF + 128 + 1024;
constexpr int AW_THREADS = 192;

struct AttnWideParams { int T, heads, C, D, new_order; float scale_log2; bf16* out; };

__device__ __forceinline__ float aw_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(AW_THREADS, 1) attn_wide_kernel(const __grid_constant__ CUtensorMap map, const AttnWideParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + AW_BAR_OFF);      // [3]
  uint64_t* empty = full + AW_STAGES;                    // [3]
  uint64_t* bar_s = full + 2 * AW_STAGES;
  uint64_t* bar_p = bar_s + 1;                           // 128 arrivals: P written, S no longer needed
  uint64_t* bar_o = bar_s + 2;
  uint32_t* tmem_slot = (uint32_t*)(bar_s + 3);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = blockIdx.x, b = blockIdx.z, h = blockIdx.y;      // grid (query tile, head, sample): no prologue division
  const int D = p.D, n_c = D >> 6;
  const int qcol = p.new_order ? h * D : h * 3 * D;
  const int kcol = p.new_order ? p.C + h * D : h * 3 * D + D;
  const int vcol = p.new_order ? 2 * p.C + h * D : h * 3 * D + 2 * D;

  pdl_launch_dependents();
  if (tid == 0) {
    prefetch_tmap(&map);
    for (int s = 0; s < AW_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(bar_s, 1); mbar_init(bar_p, 128); mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int c = 0; c < n_c; ++c) {                    // (Q_c, K_c): 16 KB + 32 KB
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], 3 * 16384);
        uint8_t* sp = smem + stage * AW_STAGE_BYTES;
        tma_load_3d(sp, &map, &full[stage], qcol + 64 * c, mt * AW_M, b);
        tma_load_3d(sp + 16384, &map, &full[stage], kcol + 64 * c, 0, b);
        tma_load_3d(sp + 32768, &map, &full[stage], kcol + 64 * c, 128, b);
        if (++stage == AW_STAGES) { stage = 0; phase ^= 1; }
      }
      for (int c = 0; c < n_c; ++c) {                    // V piece c: 256 keys x 64 channels
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], 2 * 16384);
        uint8_t* sp = smem + stage * AW_STAGE_BYTES;
        tma_load_3d(sp, &map, &full[stage], vcol + 64 * c, 0, b);
        tma_load_3d(sp + 16384, &map, &full[stage], vcol + 64 * c, 128, b);
        if (++stage == AW_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0; uint32_t phase = 0;
    const uint32_t idesc_s = make_idesc(AW_M, AW_T);
    for (int c = 0; c < n_c; ++c) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + stage * AW_STAGE_BYTES);
        const uint64_t ad = make_desc_sw128(sa), bd = make_desc_sw128(sa + 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc_s, (c > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty[stage]);
        if (c == n_c - 1) umma_commit(bar_s);
      }
      __syncwarp();
      if (++stage == AW_STAGES) { stage = 0; phase ^= 1; }
    }
    mbar_wait(bar_p, 0);                                  // P complete; every softmax thread is done with S
    tc_fence_after();
    const uint32_t idesc_o = make_idesc_major(AW_M, 64, 0, 1);
    const uint32_t pbase = smem_u32(smem + AW_P_OFF);
    for (int c = 0; c < n_c; ++c) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t vbase = smem_u32(smem + stage * AW_STAGE_BYTES);
#pragma unroll
        for (int j = 0; j < AW_T / 16; ++j) {
          const uint64_t ad = make_desc_sw128(pbase + (j >> 2) * 16384 + (j & 3) * 32);
          const uint64_t bd = make_desc_sw128_mn(vbase + j * 2048, 1024);
          umma_bf16(tmem + (uint32_t)(64 * c), ad, bd, idesc_o, j > 0);
        }
        umma_commit(&empty[stage]);
        if (c == n_c - 1) umma_commit(bar_o);
      }
      __syncwarp();
      if (++stage == AW_STAGES) { stage = 0; phase ^= 1; }
    }
  } else {
    // ===================== softmax + epilogue (warps 2..5: TMEM lane quarter warp & 3) =====================
    const int r = (warp & 3) * 32 + lane;              // query row of this thread = TMEM lane
    const uint32_t t_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    mbar_wait(bar_s, 0);
    tc_fence_after();
    float mx = -INFINITY;
    for (int c0 = 0; c0 < AW_T; c0 += 32) {
      if (c0 >= p.T) break;
      uint32_t v[32];
      tmem_ld32(t_row + (uint32_t)c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) if (c0 + j < p.T) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    const float mxs = mx * p.scale_log2;
    float sum = 0.f;
    uint8_t* prow = smem + AW_P_OFF + r * 128;
    for (int c0 = 0; c0 < AW_T; c0 += 32) {
      uint32_t v[32];
      if (c0 < p.T) { tmem_ld32(t_row + (uint32_t)c0, v); tmem_ld_wait(); }
      uint8_t* pchunk = prow + (c0 >> 6) * 16384;
      const int c16 = (c0 & 63) >> 3;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 o4;
        __nv_bfloat162* o2 = (__nv_bfloat162*)&o4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = c0 + i * 8 + 2 * q;
          const float e0 = c < p.T ? aw_ex2(fmaf(__uint_as_float(v[i * 8 + 2 * q]), p.scale_log2, -mxs)) : 0.f;
          const float e1 = c + 1 < p.T ? aw_ex2(fmaf(__uint_as_float(v[i * 8 + 2 * q + 1]), p.scale_log2, -mxs)) : 0.f;
          sum += e0 + e1;
          o2[q] = __floats2bfloat162_rn(e0, e1);
        }
        sts_u4(smem_u32(pchunk) + (uint32_t)(((c16 + i) ^ (r & 7)) << 4), o4);     // explicit st.shared (a generic store resolves the space at run time)
      }
    }
    fence_proxy_async();          // P was written through the generic proxy; the MMA reads it through the async proxy
    tc_fence_before();
    mbar_arrive(bar_p);
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float inv = 1.0f / sum;
    {
      // tcgen05.ld is warp-collective (.sync.aligned): every lane loads, only rows that exist store
      const bool row_ok = mt * AW_M + r < p.T;
      bf16* op = p.out + ((long long)b * p.T + mt * AW_M + (row_ok ? r : 0)) * p.C + h * D;
      for (int c0 = 0; c0 < D; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(t_row + (uint32_t)c0, v);
        tmem_ld_wait();
        uint4 o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          __nv_bfloat162* o2 = (__nv_bfloat162*)&o[i];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            o2[q] = __floats2bfloat162_rn(__uint_as_float(v[i * 8 + 2 * q]) * inv, __uint_as_float(v[i * 8 + 2 * q + 1]) * inv);
        }
        if (row_ok) {
          stg256(op + c0, o[0], o[1]);
          stg256(op + c0 + 16, o[2], o[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------------------------------------
struct AttnWidePlan { std::map<int, CUtensorMap> maps; };
static std::map<const Op*, AttnWidePlan> g_wide_plans;   // keyed by op address (ops vector is stable after build)
static std::mutex g_wide_mu;   // engines on different host threads share the map (each touches only its own ops' entries)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_wide_encode = nullptr;

bool attn_wide_supported(const Engine& e, const Op& op) {
  if (!e.bf16 || op.kind != OP_ATTN) return false;
  const char* off = tuning_env("CFM_DISABLE_WIDE_ATTN");
  if (off && off[0] == '1') return false;
  return op.ch > 64 && op.ch <= 512 && op.ch % 64 == 0 && op.Hin * op.Win <= AW_T;
}

int attn_wide_launch(Engine& e, const Op& op, int B, cudaStream_t st) {
  if (!g_wide_encode) {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { e.err = "cuTensorMapEncodeTiled unavailable"; return CFM_ERR_CUDA; }
    g_wide_encode = (EncodeTiledFn)fn;
  }
  static DeviceOnce attr;
  if (attr.pending(e.device)) {
    if (cudaFuncSetAttribute(attn_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AW_SMEM) != cudaSuccess) { e.err = "cudaFuncSetAttribute(attn_wide_kernel) failed"; return CFM_ERR_CUDA; }
    attr.done(e.device);
  }
  const int T = op.Hin * op.Win;
  AttnWidePlan* plp;
  { std::lock_guard<std::mutex> lk(g_wide_mu); plp = &g_wide_plans[&op]; }
  AttnWidePlan& pl = *plp;
  const void* qkv = tensor_ptr(e, op.src0, B);
  auto it = pl.maps.find(B);
  if (it == pl.maps.end()) {
    CUtensorMap m;
    const int C3 = 3 * op.Cin;
    cuuint64_t dims[3] = {(cuuint64_t)C3, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)C3 * 2, (cuuint64_t)T * C3 * 2};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_wide_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)qkv, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { e.err = "cuTensorMapEncodeTiled(qkv, wide) failed"; return CFM_ERR_CUDA; }
    it = pl.maps.emplace(B, m).first;
  }
  AttnWideParams p{};
  p.T = T; p.heads = op.heads; p.C = op.Cin; p.D = op.ch; p.new_order = e.cfg.use_new_attention_order;
  p.scale_log2 = (1.0f / sqrtf((float)op.ch)) * 1.4426950408889634f;
  p.out = (bf16*)tensor_ptr(e, op.out, B);
  if (B > 65535) { e.err = "attn_wide: batch too large for the grid"; return CFM_ERR_INVALID; }
  LaunchCfg lc(dim3((T + AW_M - 1) / AW_M, op.heads, B), dim3(AW_THREADS), AW_SMEM, st, 1, pdl_enabled());
  cudaError_t ce = cudaLaunchKernelEx(&lc.cfg, attn_wide_kernel, it->second, p);
  if (ce != cudaSuccess) { e.err = std::string("attn_wide_kernel launch failed: ") + cudaGetErrorString(ce); return CFM_ERR_CUDA; }
  return 0;
}

void attn_wide_release(Engine& e) {
  std::lock_guard<std::mutex> lk(g_wide_mu);
  for (const Op& op : e.ops) {
    auto it = g_wide_plans.find(&op);
    if (it != g_wide_plans.end()) it->second.maps.clear();
  }
}

void attn_wide_forget(Engine& e) {
  std::lock_guard<std::mutex> lk(g_wide_mu);
  for (const Op& op : e.ops) g_wide_plans.erase(&op);
}

}  // namespace cfm
