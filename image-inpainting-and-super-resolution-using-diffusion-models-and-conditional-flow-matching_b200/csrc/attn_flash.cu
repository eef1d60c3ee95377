// Attention core for any sequence length on tcgen05/TMEM (AttentionBlock of unet.py:424-483 at shapes the
// fixed T = 256 kernel of attn_tc.cu does not take: MNIST 28x28 -> T = 784 with one 32-wide head, 7x7 -> T = 49, ...).
//
//   a = softmax((q s)^T (k s)) v,  s = ch^-1/4, per (sample, head); head width D = 32 or 64.
//
// One CTA per (sample, head, 128-query tile); the keys are walked in blocks of 128 with an online softmax:
//   S_j = Q K_j^T          one UMMA chain (128 x 128 x D) into TMEM
//   thread r owns query row r: block max / running max, p = 2^(s - m), running sum; P_j (bf16) goes to shared
//   memory in the swizzled K-major layout                                   (two tcgen05.ld passes over S_j)
//   O_j = P_j V_j          second UMMA chain (V consumed as an MN-major operand straight from its NHWC rows)
//   o = o * 2^(m_old - m_new) + O_j   in registers (D fp32 per thread)
// K_{j+1} / V_{j+1} arrive by TMA (3-D map {3C, T, B}: rows past T are zero-filled, and masked) while block j is
// processed; S_{j+1} is issued together with O_j so the tensor pipe overlaps the softmax.  No atomics; a sample's
// result does not depend on its batch.
#include <cstring>
#include <map>
#include <mutex>
#include "engine.h"
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cfm {

void* tensor_ptr(const Engine& e, int id, int B);

constexpr int AF_M = 128, AF_BK = 128;
constexpr int AF_Q_OFF = 0, AF_K_OFF = 16384, AF_V_OFF = 49152, AF_P_OFF = 81920, AF_BAR_OFF = 114688;
constexpr int AF_XCH_OFF = AF_BAR_OFF + 128;
constexpr int AF_SMEM = AF_XCH_OFF + 2 * 256 * 4 + 1024;

struct AttnFlashParams { int T, heads, C, new_order, n_kv, pack, B; float scale_log2; bf16* out; };   // pack: samples per 128-row tile (T <= 64)

__device__ __forceinline__ float af_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 64 scores of one row (this thread's half of a 128-key block, in registers) -> p = 2^(s*scale - m), partial row sum,
// and P as bf16 in shared memory (K-major SWIZZLE_128B; a thread's 64 keys are exactly one 64-key sub-tile row).
// kMasked: only keys in [lo, hi) belong to this row (last block of a long sequence, or the row's own sample when
// several short sequences share a tile).  Independent sum chains keep dependent-issue latency off the critical path.
template <bool kMasked>
__device__ __forceinline__ float af_local_max(const uint32_t (&sv)[2][32], int kbase, int lo, int hi) {
  float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (!kMasked || (kbase + c * 32 + i >= lo && kbase + c * 32 + i < hi)) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(sv[c][i]));
  return fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
}
template <bool kMasked>
__device__ __forceinline__ float af_exp_block(const uint32_t (&sv)[2][32], int kbase, int lo, int hi, float scale_log2, float m_new,
                                              uint8_t* prow64, int r) {
  float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int c16 = c * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 o4;
      __nv_bfloat162* o2 = (__nv_bfloat162*)&o4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int cc = kbase + c * 32 + i * 8 + 2 * q;
        float e0 = af_ex2(fmaf(__uint_as_float(sv[c][i * 8 + 2 * q]), scale_log2, -m_new));
        float e1 = af_ex2(fmaf(__uint_as_float(sv[c][i * 8 + 2 * q + 1]), scale_log2, -m_new));
        if (kMasked) { if (cc < lo || cc >= hi) e0 = 0.f; if (cc + 1 < lo || cc + 1 >= hi) e1 = 0.f; }
        s4[q] += e0 + e1;
        o2[q] = __floats2bfloat162_rn(e0, e1);
      }
      sts_u4(smem_u32(prow64) + (uint32_t)(((c16 + i) ^ (r & 7)) << 4), o4);     // explicit st.shared (a generic store resolves the space at run time)
    }
  }
  return (s4[0] + s4[1]) + (s4[2] + s4[3]);
}

// 256 threads: thread t and t + 128 share query row (t & 127) - warps w and w + 4 may both read TMEM lanes
// 32 (w & 3) .. +31.  Each takes 64 of a block's 128 keys (scores in registers: one TMEM pass) and D/2 of the output
// channels; the running max is exchanged through shared memory once per block, the row sums only at the end.
// 16 warps per SM (2 CTAs) instead of 8 hide the MUFU / TMEM latencies of this exp-bound loop.
template <int D>
__global__ void __launch_bounds__(256, 2) attn_flash_kernel(const __grid_constant__ CUtensorMap map, const AttnFlashParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar_q = (uint64_t*)(smem + AF_BAR_OFF);
  uint64_t* bar_k = bar_q + 1;      // [2]
  uint64_t* bar_v = bar_q + 3;      // [2]
  uint64_t* bar_s = bar_q + 5;
  uint64_t* bar_o = bar_q + 6;
  uint32_t* tmem_slot = (uint32_t*)(bar_q + 7);
  float* xch = (float*)(smem + AF_XCH_OFF);            // [2][256]: per-block max (double-buffered by block parity)
  const int tid = threadIdx.x, warp = tid >> 5;
  const int mt = blockIdx.x, b = (int)blockIdx.z * p.pack, h = blockIdx.y;   // b: first sample of the tile; grid (tile, head, sample group): no prologue division
  const int qcol = p.new_order ? h * D : h * 3 * D;
  const int kcol = p.new_order ? p.C + h * D : h * 3 * D + D;
  const int vcol = p.new_order ? 2 * p.C + h * D : h * 3 * D + 2 * D;
  constexpr int ROWB = D * 2;                 // bytes of one q / k / v row
  constexpr int DH = D / 2;                   // output channels per thread
  const int TILE_B = (p.pack > 1 ? p.pack * p.T : 128) * ROWB;   // bytes of one operand tile (box rows x row bytes)

  pdl_launch_dependents();
  if (tid == 0) {
    prefetch_tmap(&map);
    for (int i = 0; i < 7; ++i) mbar_init(bar_q + i, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_s = tmem, tmem_o = tmem + 128;

  const uint32_t idesc_s = make_idesc(AF_M, AF_BK);
  const uint32_t idesc_o = make_idesc_major(AF_M, D, 0, 1);
  const uint32_t q_addr = smem_u32(smem + AF_Q_OFF), p_addr = smem_u32(smem + AF_P_OFF);

  // S_j = Q K_j^T (one elected thread)
  auto issue_s = [&](int j) {
    const uint64_t ad = make_desc_k(q_addr, D), bd = make_desc_k(smem_u32(smem + AF_K_OFF + (j & 1) * 16384), D);
#pragma unroll
    for (int k = 0; k < D / 16; ++k) umma_bf16(tmem_s, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc_s, k > 0);
    umma_commit(bar_s);
  };
  auto load_kv = [&](int j) {
    const int buf = j & 1;
    mbar_expect_tx(&bar_k[buf], TILE_B);
    tma_load_3d(smem + AF_K_OFF + buf * 16384, &map, &bar_k[buf], kcol, j * AF_BK, b);
    mbar_expect_tx(&bar_v[buf], TILE_B);
    tma_load_3d(smem + AF_V_OFF + buf * 16384, &map, &bar_v[buf], vcol, j * AF_BK, b);
  };

  if (p.pack > 1 && p.pack * p.T < 128) {
    // packed tiles that do not fill 128 rows (7x7 maps: 2 x 49): the TMA box leaves V rows [pack*T, 128) untouched, and
    // 0 (P) x stale NaN (V) would poison the PV product - clear them (visible to the MMA after fence.proxy.async below)
    uint4* vz = (uint4*)(smem + AF_V_OFF + p.pack * p.T * ROWB);
    const int nz = (128 - p.pack * p.T) * ROWB / 16;
    for (int i = tid; i < nz; i += 256) vz[i] = make_uint4(0, 0, 0, 0);
  }
  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bar_q, TILE_B);
      tma_load_3d(smem + AF_Q_OFF, &map, bar_q, qcol, mt * AF_M, b);
      load_kv(0);
      if (p.n_kv > 1) load_kv(1);
    }
    mbar_wait(bar_q, 0);
    mbar_wait(&bar_k[0], 0);
    tc_fence_after();
    if (elect_one()) issue_s(0);
    __syncwarp();
  }

  const int r = tid & 127;                     // query row of this thread = TMEM lane
  const int half = tid >> 7;                   // keys [64 half, +64) of every block, output channels [DH half, +DH)
  const uint32_t t_row = ((uint32_t)((warp & 3) * 32) << 16);
  const int kbase = half * 64;
  float m_run = -INFINITY, l_run = 0.f;        // running max (log2 units, shared by the pair) and this thread's partial sum
  float o[DH];
#pragma unroll
  for (int i = 0; i < DH; ++i) o[i] = 0.f;
  uint8_t* prow64 = smem + AF_P_OFF + half * 16384 + r * 128;

  for (int j = 0; j < p.n_kv; ++j) {
    const int kv0 = j * AF_BK;
    // keys of this block that belong to this row: the existing ones, or (packed tiles) those of the row's own sample
    int lo = 0, hi = min(AF_BK, p.T - kv0);
    if (p.pack > 1) { const int sidx = min(r / p.T, p.pack - 1); lo = sidx * p.T; hi = lo + p.T; }
    const bool masked = !(lo == 0 && hi == AF_BK);
    mbar_wait(bar_s, j & 1);
    tc_fence_after();
    uint32_t sv[2][32];
    tmem_ld32(tmem_s + t_row + (uint32_t)kbase, sv[0]);
    tmem_ld32(tmem_s + t_row + (uint32_t)(kbase + 32), sv[1]);
    tmem_ld_wait();
    const float lmx = masked ? af_local_max<true>(sv, kbase, lo, hi) : af_local_max<false>(sv, kbase, lo, hi);
    float* xb = xch + (j & 1) * 256;
    xb[tid] = lmx;
    __syncthreads();
    const float mx = fmaxf(lmx, xb[tid ^ 128]);
    const float m_new = fmaxf(m_run, mx * p.scale_log2);
    const float alpha = af_ex2(m_run - m_new);           // 0 for the first block (m_run = -inf)
    m_run = m_new;
    const float sum = masked ? af_exp_block<true>(sv, kbase, lo, hi, p.scale_log2, m_new, prow64, r)
                             : af_exp_block<false>(sv, kbase, lo, hi, p.scale_log2, m_new, prow64, r);
    l_run = l_run * alpha + sum;
    fence_proxy_async();          // P was written through the generic proxy; the MMA reads it through the async proxy
    tc_fence_before();
    __syncthreads();              // every row of P written, every thread done reading S_j
    if (warp == 0) {
      mbar_wait(&bar_v[j & 1], (j >> 1) & 1);
      if (j + 1 < p.n_kv) mbar_wait(&bar_k[(j + 1) & 1], ((j + 1) >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t vbase = smem_u32(smem + AF_V_OFF + (j & 1) * 16384);
#pragma unroll
        for (int k = 0; k < AF_BK / 16; ++k) {
          const uint64_t ad = make_desc_k(p_addr + (k >> 2) * 16384 + (k & 3) * 32, 64);
          const uint64_t bd = make_desc_mn(vbase + k * 16 * ROWB, 1024, D);
          umma_bf16(tmem_o, ad, bd, idesc_o, k > 0);
        }
        umma_commit(bar_o);
        if (j + 1 < p.n_kv) issue_s(j + 1);   // overlaps the o-update below and the next block's TMA wait
      }
      __syncwarp();
    }
    // ---- o = o * alpha + O_j (this thread's DH channels) ----
    mbar_wait(bar_o, j & 1);
    tc_fence_after();
    {
      uint32_t v[DH];
      if (DH == 32) tmem_ld32(tmem_o + t_row + (uint32_t)(half * DH), (uint32_t*)v);
      else tmem_ld16(tmem_o + t_row + (uint32_t)(half * DH), (uint32_t*)v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < DH; ++i) o[i] = fmaf(o[i], alpha, __uint_as_float(v[i]));
    }
    // K_j / V_j buffers are free (S_j and O_j have completed): fetch block j + 2 into them
    if (warp == 0 && j + 2 < p.n_kv) {
      if (elect_one()) load_kv(j + 2);
      __syncwarp();
    }
    tc_fence_before();            // order this thread's TMEM reads before the next block's MMAs (issued after the next sync)
  }

  // ---- combine the pair's partial row sums, normalise, store this thread's DH channels with 256-bit stores ----
  __syncthreads();
  xch[tid] = l_run;
  __syncthreads();
  const float inv = 1.0f / (l_run + xch[tid ^ 128]);
  // row -> (sample, token): one sample per tile, or p.pack short sequences back to back
  const int s_idx = p.pack > 1 ? r / p.T : 0;
  const int tok = p.pack > 1 ? r - s_idx * p.T : mt * AF_M + r;
  if (tok < p.T && s_idx < p.pack && b + s_idx < p.B) {
    bf16* op = p.out + ((long long)(b + s_idx) * p.T + tok) * p.C + h * D + half * DH;
#pragma unroll
    for (int c0 = 0; c0 < DH; c0 += 16) {
      uint4 ov[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        __nv_bfloat162* o2 = (__nv_bfloat162*)&ov[i];
#pragma unroll
        for (int q = 0; q < 4; ++q) o2[q] = __floats2bfloat162_rn(o[c0 + i * 8 + 2 * q] * inv, o[c0 + i * 8 + 2 * q + 1] * inv);
      }
      stg256(op + c0, ov[0], ov[1]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

// ------------------------------------------------------------------------------------------------
// Long sequences (T > 128: the 28x28 MNIST maps): the same online-softmax loop with 64-key blocks and one thread per
// query row.  A CTA then needs 128 TMEM columns (S 64 + O <= 64) and 42 / 66 KB of shared memory, so 4 (D = 32) or
// 3 (D = 64) CTAs share an SM instead of 2: the loop is a chain of short dependent phases (MMA round trip, TMEM load,
// exp, shared-memory store, barrier) and only more independent CTAs keep the MUFU / ALU pipes busy.
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128, D == 32 ? 4 : 3)
attn_flash64_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapKV, const AttnFlashParams p) {
  constexpr int BK = 64;
  constexpr int ROWB = D * 2, QB = 128 * ROWB, KB = BK * ROWB;
  constexpr int K_OFF = QB, V_OFF = QB + 2 * KB, P_OFF = QB + 4 * KB, BAR_OFF = P_OFF + 16384;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar_q = (uint64_t*)(smem + BAR_OFF);
  uint64_t* bar_k = bar_q + 1;      // [2]
  uint64_t* bar_v = bar_q + 3;      // [2]
  uint64_t* bar_s = bar_q + 5;
  uint64_t* bar_o = bar_q + 6;
  uint32_t* tmem_slot = (uint32_t*)(bar_q + 7);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int mt = blockIdx.x, b = blockIdx.z, h = blockIdx.y;      // grid (query tile, head, sample): no prologue division
  const int qcol = p.new_order ? h * D : h * 3 * D;
  const int kcol = p.new_order ? p.C + h * D : h * 3 * D + D;
  const int vcol = p.new_order ? 2 * p.C + h * D : h * 3 * D + 2 * D;
  const int n_kv = (p.T + BK - 1) / BK;

  pdl_launch_dependents();
  if (tid == 0) {
    prefetch_tmap(&mapQ); prefetch_tmap(&mapKV);
    for (int i = 0; i < 7; ++i) mbar_init(bar_q + i, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_s = tmem, tmem_o = tmem + 64;
  const uint32_t idesc_s = make_idesc(128, BK);
  const uint32_t idesc_o = make_idesc_major(128, D, 0, 1);
  const uint32_t q_addr = smem_u32(smem), p_addr = smem_u32(smem + P_OFF);

  auto issue_s = [&](int j) {
    const uint64_t ad = make_desc_k(q_addr, D), bd = make_desc_k(smem_u32(smem + K_OFF + (j & 1) * KB), D);
#pragma unroll
    for (int k = 0; k < D / 16; ++k) umma_bf16(tmem_s, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc_s, k > 0);
    umma_commit(bar_s);
  };
  auto load_kv = [&](int j) {
    const int buf = j & 1;
    mbar_expect_tx(&bar_k[buf], KB);
    tma_load_3d(smem + K_OFF + buf * KB, &mapKV, &bar_k[buf], kcol, j * BK, b);
    mbar_expect_tx(&bar_v[buf], KB);
    tma_load_3d(smem + V_OFF + buf * KB, &mapKV, &bar_v[buf], vcol, j * BK, b);
  };

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bar_q, QB);
      tma_load_3d(smem, &mapQ, bar_q, qcol, mt * 128, b);
      load_kv(0);
      if (n_kv > 1) load_kv(1);
    }
    mbar_wait(bar_q, 0);
    mbar_wait(&bar_k[0], 0);
    tc_fence_after();
    if (elect_one()) issue_s(0);
    __syncwarp();
  }

  const int r = tid;                           // query row of this thread = TMEM lane
  const uint32_t t_row = ((uint32_t)(warp * 32) << 16);
  float m_run = -INFINITY, l_run = 0.f;
  float o[D];
#pragma unroll
  for (int i = 0; i < D; ++i) o[i] = 0.f;
  uint8_t* prow = smem + P_OFF + r * 128;

  for (int j = 0; j < n_kv; ++j) {
    const int hi = min(BK, p.T - j * BK);      // keys of this block that exist
    const bool masked = hi != BK;
    mbar_wait(bar_s, j & 1);
    tc_fence_after();
    uint32_t sv[2][32];
    tmem_ld32(tmem_s + t_row, sv[0]);
    tmem_ld32(tmem_s + t_row + 32u, sv[1]);
    tmem_ld_wait();
    const float mx = masked ? af_local_max<true>(sv, 0, 0, hi) : af_local_max<false>(sv, 0, 0, hi);
    const float m_new = fmaxf(m_run, mx * p.scale_log2);
    const float alpha = af_ex2(m_run - m_new);           // 0 for the first block (m_run = -inf)
    m_run = m_new;
    const float sum = masked ? af_exp_block<true>(sv, 0, 0, hi, p.scale_log2, m_new, prow, r)
                             : af_exp_block<false>(sv, 0, 0, hi, p.scale_log2, m_new, prow, r);
    l_run = l_run * alpha + sum;
    fence_proxy_async();          // P was written through the generic proxy; the MMA reads it through the async proxy
    tc_fence_before();
    __syncthreads();              // every row of P written, every thread done reading S_j
    if (warp == 0) {
      mbar_wait(&bar_v[j & 1], (j >> 1) & 1);
      if (j + 1 < n_kv) mbar_wait(&bar_k[(j + 1) & 1], ((j + 1) >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t vbase = smem_u32(smem + V_OFF + (j & 1) * KB);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t ad = make_desc_k(p_addr + k * 32, 64);
          const uint64_t bd = make_desc_mn(vbase + k * 16 * ROWB, 1024, D);
          umma_bf16(tmem_o, ad, bd, idesc_o, k > 0);
        }
        umma_commit(bar_o);
        if (j + 1 < n_kv) issue_s(j + 1);
      }
      __syncwarp();
    }
    mbar_wait(bar_o, j & 1);
    tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < D; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_o + t_row + (uint32_t)c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[c0 + i] = fmaf(o[c0 + i], alpha, __uint_as_float(v[i]));
    }
    if (warp == 0 && j + 2 < n_kv) {             // K_j / V_j buffers are free: fetch block j + 2 into them
      if (elect_one()) load_kv(j + 2);
      __syncwarp();
    }
    tc_fence_before();
  }

  const float inv = 1.0f / l_run;
  if (mt * 128 + r < p.T) {
    bf16* op = p.out + ((long long)b * p.T + mt * 128 + r) * p.C + h * D;
#pragma unroll
    for (int c0 = 0; c0 < D; c0 += 16) {
      uint4 ov[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        __nv_bfloat162* o2 = (__nv_bfloat162*)&ov[i];
#pragma unroll
        for (int q = 0; q < 4; ++q) o2[q] = __floats2bfloat162_rn(o[c0 + i * 8 + 2 * q] * inv, o[c0 + i * 8 + 2 * q + 1] * inv);
      }
      stg256(op + c0, ov[0], ov[1]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 128); }
}

// ------------------------------------------------------------------------------------------------
struct AttnFlashPlan { std::map<int, CUtensorMap> maps; std::map<int, CUtensorMap> maps64; };   // maps64: 64-row K / V boxes
static std::map<const Op*, AttnFlashPlan> g_flash_plans;   // keyed by op address (ops vector is stable after build)
static std::mutex g_flash_mu;   // engines on different host threads share the map (each touches only its own ops' entries)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_flash_encode = nullptr;

bool attn_flash_supported(const Engine& e, const Op& op) {
  if (!e.bf16 || op.kind != OP_ATTN) return false;
  const char* off = tuning_env("CFM_DISABLE_FLASH_ATTN");
  if (off && off[0] == '1') return false;
  return (op.ch == 32 || op.ch == 64) && op.Cin % 8 == 0;
}

int attn_flash_launch(Engine& e, const Op& op, int B, cudaStream_t st) {
  if (!g_flash_encode) {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { e.err = "cuTensorMapEncodeTiled unavailable"; return CFM_ERR_CUDA; }
    g_flash_encode = (EncodeTiledFn)fn;
  }
  static DeviceOnce attr;
  if (attr.pending(e.device)) {
    if (cudaFuncSetAttribute(attn_flash_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, AF_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(attn_flash_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AF_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(attn_flash64_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(attn_flash64_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess) {
      e.err = "cudaFuncSetAttribute(attn_flash_kernel) failed"; return CFM_ERR_CUDA;
    }
    attr.done(e.device);
  }
  const int T = op.Hin * op.Win;
  AttnFlashPlan* plp;
  { std::lock_guard<std::mutex> lk(g_flash_mu); plp = &g_flash_plans[&op]; }
  AttnFlashPlan& pl = *plp;
  const void* qkv = tensor_ptr(e, op.src0, B);
  auto it = pl.maps.find(B);
  if (it == pl.maps.end()) {
    CUtensorMap m;
    const int C3 = 3 * op.Cin;
    cuuint64_t dims[3] = {(cuuint64_t)C3, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)C3 * 2, (cuuint64_t)T * C3 * 2};
    const int pack = T <= 64 ? 128 / T : 1;
    cuuint32_t box[3] = {(cuuint32_t)op.ch, (cuuint32_t)(pack > 1 ? T : 128), (cuuint32_t)pack};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_flash_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)qkv, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                op.ch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { e.err = "cuTensorMapEncodeTiled(qkv, flash) failed"; return CFM_ERR_CUDA; }
    it = pl.maps.emplace(B, m).first;
  }
  AttnFlashParams p{};
  p.T = T; p.heads = op.heads; p.C = op.Cin; p.new_order = e.cfg.use_new_attention_order;
  p.n_kv = (T + AF_BK - 1) / AF_BK;
  p.pack = T <= 64 ? 128 / T : 1;      // short sequences (middle blocks: 4x4, 7x7, 8x8 maps): several samples per tile
  p.B = B;
  p.scale_log2 = (1.0f / sqrtf((float)op.ch)) * 1.4426950408889634f;
  p.out = (bf16*)tensor_ptr(e, op.out, B);
  static const bool no64 = [] { const char* v = tuning_env("CFM_DISABLE_FLASH64"); return v && v[0] == '1'; }();
  if (T > 128 && !no64) {
    // long sequences: 64-key blocks, 3-4 CTAs per SM
    auto it64 = pl.maps64.find(B);
    if (it64 == pl.maps64.end()) {
      CUtensorMap m;
      const int C3 = 3 * op.Cin;
      cuuint64_t dims[3] = {(cuuint64_t)C3, (cuuint64_t)T, (cuuint64_t)B};
      cuuint64_t strides[2] = {(cuuint64_t)C3 * 2, (cuuint64_t)T * C3 * 2};
      cuuint32_t box[3] = {(cuuint32_t)op.ch, 64, 1};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = g_flash_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)qkv, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  op.ch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { e.err = "cuTensorMapEncodeTiled(qkv, flash64) failed"; return CFM_ERR_CUDA; }
      it64 = pl.maps64.emplace(B, m).first;
    }
    const int rowb = op.ch * 2;
    const size_t smem64 = (size_t)128 * rowb + 4 * 64 * rowb + 16384 + 128 + 1024;
    if (B > 65535) { e.err = "attn_flash64: batch too large for the grid"; return CFM_ERR_INVALID; }
    LaunchCfg lc64(dim3((T + 127) / 128, op.heads, B), dim3(128), smem64, st, 1, pdl_enabled());
    cudaError_t ce64 = op.ch == 64 ? cudaLaunchKernelEx(&lc64.cfg, attn_flash64_kernel<64>, it->second, it64->second, p)
                                   : cudaLaunchKernelEx(&lc64.cfg, attn_flash64_kernel<32>, it->second, it64->second, p);
    if (ce64 != cudaSuccess) { e.err = std::string("attn_flash64_kernel launch failed: ") + cudaGetErrorString(ce64); return CFM_ERR_CUDA; }
    return 0;
  }
  if ((B + p.pack - 1) / p.pack > 65535) { e.err = "attn_flash: batch too large for the grid"; return CFM_ERR_INVALID; }
  LaunchCfg lc(dim3((T + AF_M - 1) / AF_M, op.heads, (B + p.pack - 1) / p.pack), dim3(256), AF_SMEM, st, 1, pdl_enabled());
  cudaError_t ce = op.ch == 64 ? cudaLaunchKernelEx(&lc.cfg, attn_flash_kernel<64>, it->second, p)
                               : cudaLaunchKernelEx(&lc.cfg, attn_flash_kernel<32>, it->second, p);
  if (ce != cudaSuccess) { e.err = std::string("attn_flash_kernel launch failed: ") + cudaGetErrorString(ce); return CFM_ERR_CUDA; }
  return 0;
}

void attn_flash_release(Engine& e) {
  std::lock_guard<std::mutex> lk(g_flash_mu);
  for (const Op& op : e.ops) {
    auto it = g_flash_plans.find(&op);
    if (it != g_flash_plans.end()) { it->second.maps.clear(); it->second.maps64.clear(); }
  }
}

void attn_flash_forget(Engine& e) {
  std::lock_guard<std::mutex> lk(g_flash_mu);
  for (const Op& op : e.ops) g_flash_plans.erase(&op);
}

}  // namespace cfm
