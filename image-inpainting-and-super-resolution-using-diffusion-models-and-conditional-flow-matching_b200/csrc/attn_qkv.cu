// AttentionBlock core with the qkv projection fused in (unet.py:395-401, 433-448), for the CIFAR shape:
// T = 256 tokens, 64-wide heads, legacy qkv ordering.  One CTA per (sample, head):
//
//   A  [q|k|v] = X W_h^T            X = GroupNorm output [256 tokens x C], W_h = the head's 192 rows of the 1x1 qkv conv;
//                                    64-channel chunks of X and W_h through a 2-stage TMA ring, UMMA 128 x 192 x 16 on
//                                    both 128-token halves -> TMEM columns [0, 384)
//   B  + bias -> bf16 Q, K, V in shared memory (swizzled K-major rows, exactly what TMA would have written)
//   C  S = Q K^T for both query halves -> TMEM columns [0, 512)  (the projection accumulators are dead by then)
//   D  exact two-pass fp32 softmax, one thread per query row -> bf16 P in shared memory (over the dead Q / K / ring)
//   E  O = P V (V as an MN-major operand) -> TMEM;  O / rowsum -> bf16 -> global
//
// The separate qkv conv writes a [B, T, 3C] tensor and the attention kernel reads it back (0.8 GB per block at batch
// 1024); here q, k, v never leave the SM.  Warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = one thread per token.
// EXPERIMENTAL / opt-in: parity-green but 10 % slower than the split kernels (see attn_qkv_shape_ok).
#include <cstring>
#include <map>
#include "engine.h"
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cfm {

void* tensor_ptr(const Engine& e, int id, int B);

constexpr int AQ_T = 256, AQ_D = 64;
constexpr int AQ_V_OFF = 0, AQ_Q_OFF = 32768, AQ_K_OFF = 65536, AQ_RING_OFF = 98304;
constexpr int AQ_STAGE_BYTES = 57344;                 // X chunk 256 x 128 B + W chunk 192 x 128 B
constexpr int AQ_P_OFF = AQ_Q_OFF;                    // P [2][128 x 256] bf16 = 128 KB over Q, K and the ring
constexpr int AQ_BIAS_OFF = AQ_RING_OFF + 2 * AQ_STAGE_BYTES;      // 192 floats
constexpr int AQ_BAR_OFF = AQ_BIAS_OFF + 1024;
constexpr int AQ_SMEM = AQ_BAR_OFF + 128 + 1024;
constexpr int AQ_THREADS = 320;

struct AttnQkvParams { int heads, C, n_chunks; float scale_log2; const float* bias; bf16* out; };

__device__ __forceinline__ float aq_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(AQ_THREADS, 1)
attn_qkv_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW, const AttnQkvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* s_bias = (float*)(smem + AQ_BIAS_OFF);
  uint64_t* full = (uint64_t*)(smem + AQ_BAR_OFF);       // [2]
  uint64_t* empty = full + 2;                             // [2]
  uint64_t* bar_acc = full + 4;                           // projection accumulators complete
  uint64_t* bar_qkv = full + 5;                           // 256 arrivals: Q, K, V in shared memory
  uint64_t* bar_s = full + 6;
  uint64_t* bar_p = full + 7;                             // 256 arrivals: P in shared memory, S no longer needed
  uint64_t* bar_o = full + 8;
  uint32_t* tmem_slot = (uint32_t*)(full + 9);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads;

  pdl_launch_dependents();
  if (tid == 0) {
    prefetch_tmap(&mapX); prefetch_tmap(&mapW);
    for (int s = 0; s < 2; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(bar_acc, 1); mbar_init(bar_qkv, 256); mbar_init(bar_s, 1); mbar_init(bar_p, 256); mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < 3 * AQ_D; i += AQ_THREADS) s_bias[i] = p.bias[h * 3 * AQ_D + i];   // weights: not produced by a predecessor
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int c = 0; c < p.n_chunks; ++c) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], AQ_STAGE_BYTES);
        uint8_t* sp = smem + AQ_RING_OFF + stage * AQ_STAGE_BYTES;
        tma_load_3d(sp, &mapX, &full[stage], 64 * c, 0, b);
        tma_load_3d(sp + 16384, &mapX, &full[stage], 64 * c, 128, b);
        tma_load_2d(sp + 32768, &mapW, &full[stage], 0, c * 3 * p.C + h * 3 * AQ_D);
        if (++stage == 2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0; uint32_t phase = 0;
    const uint32_t idesc_a = make_idesc(128, 3 * AQ_D);
    for (int c = 0; c < p.n_chunks; ++c) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + AQ_RING_OFF + stage * AQ_STAGE_BYTES);
        const uint64_t x0 = make_desc_sw128(sa), x1 = make_desc_sw128(sa + 16384), wd = make_desc_sw128(sa + 32768);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t acc = (c > 0 || k > 0) ? 1u : 0u;
          umma_bf16(tmem, x0 + (uint64_t)(2 * k), wd + (uint64_t)(2 * k), idesc_a, acc);
          umma_bf16(tmem + 192u, x1 + (uint64_t)(2 * k), wd + (uint64_t)(2 * k), idesc_a, acc);
        }
        umma_commit(&empty[stage]);
        if (c == p.n_chunks - 1) umma_commit(bar_acc);
      }
      __syncwarp();
      if (++stage == 2) { stage = 0; phase ^= 1; }
    }
    // ---- S = Q K^T for both query halves ----
    mbar_wait(bar_qkv, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc(128, AQ_T);
      const uint64_t kd = make_desc_sw128(smem_u32(smem + AQ_K_OFF));
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const uint64_t qd = make_desc_sw128(smem_u32(smem + AQ_Q_OFF + m * 16384));
#pragma unroll
        for (int k = 0; k < AQ_D / 16; ++k) umma_bf16(tmem + (uint32_t)(m * 256), qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc_s, k > 0);
      }
      umma_commit(bar_s);
    }
    __syncwarp();
    // ---- O = P V ----
    mbar_wait(bar_p, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t idesc_o = make_idesc_major(128, AQ_D, 0, 1);
      const uint32_t vbase = smem_u32(smem + AQ_V_OFF);
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const uint32_t pbase = smem_u32(smem + AQ_P_OFF + m * 65536);
#pragma unroll
        for (int j = 0; j < AQ_T / 16; ++j) {
          const uint64_t ad = make_desc_sw128(pbase + (j >> 2) * 16384 + (j & 3) * 32);
          const uint64_t bd = make_desc_sw128_mn(vbase + j * 2048, 1024);
          umma_bf16(tmem + (uint32_t)(m * 256), ad, bd, idesc_o, j > 0);
        }
      }
      umma_commit(bar_o);
    }
    __syncwarp();
  } else {
    // ===================== one thread per token (warps 2..9) =====================
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may read
    const int m = (warp - 2) >> 2;                      // token half: rows [128 m, +128)
    const int r = quarter * 32 + lane;                  // row within the half = TMEM lane
    const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16);
    // ---- B: projection accumulators + bias -> bf16 Q, K, V rows (swizzled like a TMA-written K-major tile) ----
    mbar_wait(bar_acc, 0);
    tc_fence_after();
#pragma unroll
    for (int ch = 0; ch < 6; ++ch) {                    // 32 columns each: q0 q1 k0 k1 v0 v1
      uint32_t v[32];
      tmem_ld32(t_lane + (uint32_t)(m * 192 + ch * 32), v);
      tmem_ld_wait();
      const int which = ch >> 1;                        // 0 q, 1 k, 2 v
      uint8_t* base = smem + (which == 0 ? AQ_Q_OFF : (which == 1 ? AQ_K_OFF : AQ_V_OFF)) + m * 16384 + r * 128;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 o4;
        __nv_bfloat162* o2 = (__nv_bfloat162*)&o4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col = ch * 32 + i * 8 + 2 * q;
          o2[q] = __floats2bfloat162_rn(__uint_as_float(v[i * 8 + 2 * q]) + s_bias[col], __uint_as_float(v[i * 8 + 2 * q + 1]) + s_bias[col + 1]);
        }
        *(uint4*)(base + ((((ch & 1) * 4 + i) ^ (r & 7)) << 4)) = o4;
      }
    }
    fence_proxy_async();          // generic-proxy writes -> visible to the MMAs (async proxy)
    tc_fence_before();
    mbar_arrive(bar_qkv);
    // ---- D: softmax of this thread's query row (256 keys) ----
    mbar_wait(bar_s, 0);
    tc_fence_after();
    const uint32_t t_s = t_lane + (uint32_t)(m * 256);
    float mx = -INFINITY;
    for (int c0 = 0; c0 < AQ_T; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(t_s + (uint32_t)c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    const float mxs = mx * p.scale_log2;
    float sum = 0.f;
    uint8_t* prow = smem + AQ_P_OFF + m * 65536 + r * 128;
    for (int c0 = 0; c0 < AQ_T; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(t_s + (uint32_t)c0, v);
      tmem_ld_wait();
      uint8_t* pchunk = prow + (c0 >> 6) * 16384;
      const int c16 = (c0 & 63) >> 3;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 o4;
        __nv_bfloat162* o2 = (__nv_bfloat162*)&o4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float e0 = aq_ex2(fmaf(__uint_as_float(v[i * 8 + 2 * q]), p.scale_log2, -mxs));
          const float e1 = aq_ex2(fmaf(__uint_as_float(v[i * 8 + 2 * q + 1]), p.scale_log2, -mxs));
          sum += e0 + e1;
          o2[q] = __floats2bfloat162_rn(e0, e1);
        }
        *(uint4*)(pchunk + (((c16 + i) ^ (r & 7)) << 4)) = o4;
      }
    }
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(bar_p);
    // ---- E: O / rowsum -> global ----
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float inv = 1.0f / sum;
    bf16* op = p.out + ((long long)b * AQ_T + m * 128 + r) * p.C + h * AQ_D;
#pragma unroll
    for (int c0 = 0; c0 < AQ_D; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(t_s + (uint32_t)c0, v);
      tmem_ld_wait();
      uint4 o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        __nv_bfloat162* o2 = (__nv_bfloat162*)&o[i];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          o2[q] = __floats2bfloat162_rn(__uint_as_float(v[i * 8 + 2 * q]) * inv, __uint_as_float(v[i * 8 + 2 * q + 1]) * inv);
      }
      stg256(op + c0, o[0], o[1]);
      stg256(op + c0 + 16, o[2], o[3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct AttnQkvPlan {
  bf16* w_packed = nullptr;      // [C / 64][3C][64]
  float* bias = nullptr;         // [3C]
  std::map<int, CUtensorMap> mapsX;
  CUtensorMap mapW; bool mapW_ok = false;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_qkv_encode = nullptr;

// shapes the fused kernel takes: 256 tokens, 64-wide heads, legacy qkv ordering, C a multiple of 64
bool attn_qkv_shape_ok(const Engine& e, int C, int heads, int T) {
  if (!e.bf16 || e.cfg.use_new_attention_order) return false;
  // Opt-in (CFM_ENABLE_FUSED_QKV=1): measured 0.368 ms per block at batch 1024 against 0.134 + 0.200 ms for the split
  // qkv conv + attention kernels.  It saves the 0.8 GB round trip, but its 512 TMEM columns and 208 KB of shared memory
  // allow one CTA per SM, so projection, softmax and PV of an item run back to back instead of overlapping.
  const char* on = getenv("CFM_ENABLE_FUSED_QKV");
  if (!(on && on[0] == '1')) return false;
  return T == AQ_T && heads > 0 && C == heads * AQ_D && C % 64 == 0 && C <= 1024;
}

int attn_qkv_prepare(Engine& e, Op& op, const float* w_oi /*[3C][C]*/, const float* bias /*[3C]*/) {
  AttnQkvPlan* pl = new AttnQkvPlan();
  const int C = op.Cin, C3 = 3 * C, n_chunks = C / 64;
  std::vector<bf16> packed((size_t)n_chunks * C3 * 64);
  for (int c = 0; c < n_chunks; ++c)
    for (int o = 0; o < C3; ++o)
      for (int j = 0; j < 64; ++j) packed[((size_t)c * C3 + o) * 64 + j] = __float2bfloat16(w_oi[(size_t)o * C + c * 64 + j]);
  void* d = nullptr;
  if (cudaMalloc(&d, packed.size() * sizeof(bf16)) != cudaSuccess) { e.err = "cudaMalloc(fused qkv weights) failed"; delete pl; return CFM_ERR_OOM; }
  e.owned.push_back(d);
  if (cudaMemcpy(d, packed.data(), packed.size() * sizeof(bf16), cudaMemcpyHostToDevice) != cudaSuccess) { e.err = "fused qkv weight upload failed"; delete pl; return CFM_ERR_CUDA; }
  pl->w_packed = (bf16*)d;
  void* bp = nullptr;
  if (cudaMalloc(&bp, sizeof(float) * C3) != cudaSuccess) { e.err = "cudaMalloc(fused qkv bias) failed"; delete pl; return CFM_ERR_OOM; }
  e.owned.push_back(bp);
  if (cudaMemcpy(bp, bias, sizeof(float) * C3, cudaMemcpyHostToDevice) != cudaSuccess) { e.err = "fused qkv bias upload failed"; delete pl; return CFM_ERR_CUDA; }
  pl->bias = (float*)bp;
  op.fq = pl;
  return 0;
}

int attn_qkv_launch(Engine& e, const Op& op, int B, cudaStream_t st) {
  AttnQkvPlan* pl = op.fq;
  if (!g_qkv_encode) {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { e.err = "cuTensorMapEncodeTiled unavailable"; return CFM_ERR_CUDA; }
    g_qkv_encode = (EncodeTiledFn)fn;
  }
  static DeviceOnce attr;
  if (attr.pending(e.device)) {
    if (cudaFuncSetAttribute(attn_qkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AQ_SMEM) != cudaSuccess) { e.err = "cudaFuncSetAttribute(attn_qkv_kernel) failed"; return CFM_ERR_CUDA; }
    attr.done(e.device);
  }
  const int C = op.Cin;
  if (!pl->mapW_ok) {
    cuuint64_t dims[2] = {64, (cuuint64_t)(C / 64) * 3 * C};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 3 * AQ_D};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_qkv_encode(&pl->mapW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pl->w_packed, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { e.err = "cuTensorMapEncodeTiled(fused qkv W) failed"; return CFM_ERR_CUDA; }
    pl->mapW_ok = true;
  }
  auto it = pl->mapsX.find(B);
  if (it == pl->mapsX.end()) {
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)AQ_T, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)AQ_T * C * 2};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_qkv_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, tensor_ptr(e, op.src0, B), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { e.err = "cuTensorMapEncodeTiled(fused qkv X) failed"; return CFM_ERR_CUDA; }
    it = pl->mapsX.emplace(B, m).first;
  }
  AttnQkvParams p{};
  p.heads = op.heads; p.C = C; p.n_chunks = C / 64;
  p.scale_log2 = (1.0f / sqrtf((float)op.ch)) * 1.4426950408889634f;
  p.bias = pl->bias;
  p.out = (bf16*)tensor_ptr(e, op.out, B);
  LaunchCfg lc(dim3(B * op.heads), dim3(AQ_THREADS), AQ_SMEM, st, 1, pdl_enabled());
  cudaError_t ce = cudaLaunchKernelEx(&lc.cfg, attn_qkv_kernel, it->second, pl->mapW, p);
  if (ce != cudaSuccess) { e.err = std::string("attn_qkv_kernel launch failed: ") + cudaGetErrorString(ce); return CFM_ERR_CUDA; }
  return 0;
}

void attn_qkv_release(Engine& e) {
  for (Op& op : e.ops)
    if (op.fq) op.fq->mapsX.clear();
}

}  // namespace cfm
