// Fused attention core on tcgen05/TMEM for the U-Net's AttentionBlock (unet.py:424-483):
//   a = softmax((q s)^T (k s)) v,  s = ch^-1/4, per (sample, head); T = 256 tokens, ch = 64.
//
// TMA brings a head's Q tiles, K and V (128-byte channel rows of the NHWC qkv tensor) into SWIZZLE_128B shared memory.
// S = Q K^T runs as UMMA chains into TMEM; each softmax thread owns one query row, reads it back with tcgen05.ld, does
// an exact two-pass softmax in fp32 and writes bf16 P into shared memory in the K-major swizzled layout; O = P V is a
// second UMMA chain (V consumed as an MN-major B operand straight from its NHWC rows) into the TMEM columns S has
// vacated; the epilogue normalises by the row sum and stores bf16.
// The shipped kernel is attn_tc_persist_kernel (one persistent CTA per SM, both query tiles of a head in flight); the
// round-1 kernel (one CTA per query tile, two CTAs per SM) is compiled only into -DCFM_TUNING builds for A/B runs.
#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include "engine.h"
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cfm {

void* tensor_ptr(const Engine& e, int id, int B);

constexpr int AT_T = 256, AT_D = 64, AT_M = 128;
#ifdef CFM_TUNING
constexpr int AT_V_OFF = 0, AT_Q_OFF = 32768, AT_K_OFF = 49152, AT_P_OFF = 32768;
constexpr int AT_BAR_OFF = 98304, AT_XCH_OFF = 98304 + 64;
constexpr int AT_SMEM = AT_XCH_OFF + 2 * 256 * 4 + 1024;
#endif

struct AttnTcParams { int heads, C, new_order; float scale_log2; bf16* out; int pf_db, pf_dh; };

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

#ifdef CFM_TUNING
// 256 threads per CTA: thread t and thread t + 128 share query row (t & 127) - warps w and w + 4 may both read TMEM
// lanes 32 (w & 3) .. +31 - and split the 256 keys (softmax) / the 64 output channels (epilogue) between them, so the
// serial load -> max -> exp -> store chain of a row is half as long and 16 warps per SM hide its latency.
__global__ void __launch_bounds__(256, 2) attn_tc_kernel(const __grid_constant__ CUtensorMap map, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar_qk = (uint64_t*)(smem + AT_BAR_OFF);
  uint64_t* bar_v = bar_qk + 1;
  uint64_t* bar_s = bar_qk + 2;
  uint64_t* bar_o = bar_qk + 3;
  uint32_t* tmem_slot = (uint32_t*)(bar_qk + 4);
  float* xch = (float*)(smem + AT_XCH_OFF);            // [2][256]: per-thread partial max / sum for the row partner
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // grid = (query tile, head, sample): no integer division in the prologue (the I2F / MUFU.RCP / F2I sequence of a
  // division queued behind the other resident CTA's exponentials: 6.7 % of the kernel's stall samples)
  const int mt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int qcol = p.new_order ? h * AT_D : h * 3 * AT_D;
  const int kcol = p.new_order ? p.C + h * AT_D : h * 3 * AT_D + AT_D;
  const int vcol = p.new_order ? 2 * p.C + h * AT_D : h * 3 * AT_D + 2 * AT_D;
  const int row0 = b * AT_T;

  pdl_launch_dependents();
  if (tid == 0) {
    prefetch_tmap(&map);
    mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);      // in parallel with warp 0's barrier initialisation
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bar_qk, 3 * 16384);
      tma_load_2d(smem + AT_Q_OFF, &map, bar_qk, qcol, row0 + mt * AT_M);
      tma_load_2d(smem + AT_K_OFF, &map, bar_qk, kcol, row0);
      tma_load_2d(smem + AT_K_OFF + 16384, &map, bar_qk, kcol, row0 + 128);
      mbar_expect_tx(bar_v, 2 * 16384);
      tma_load_2d(smem + AT_V_OFF, &map, bar_v, vcol, row0);
      tma_load_2d(smem + AT_V_OFF + 16384, &map, bar_v, vcol, row0 + 128);
      // the CTA that inherits this SM slot is pf_ahead (sample, head) pairs further on: have its q / k / v boxes in L2
      // by the time it starts (this kernel is a chain of dependent phases per CTA; its load is pure exposed latency)
      int h2 = h + p.pf_dh, b2 = b + p.pf_db;
      if (h2 >= p.heads) { h2 -= p.heads; ++b2; }
      if ((p.pf_db | p.pf_dh) != 0 && b2 < (int)gridDim.z) {
        const int q2 = p.new_order ? h2 * AT_D : h2 * 3 * AT_D;
        const int k2 = p.new_order ? p.C + h2 * AT_D : h2 * 3 * AT_D + AT_D;
        const int v2 = p.new_order ? 2 * p.C + h2 * AT_D : h2 * 3 * AT_D + 2 * AT_D;
        tma_prefetch_2d(&map, q2, b2 * AT_T + mt * AT_M);
        if (mt == 0) {                          // k and v are shared by the two query tiles of a head
          tma_prefetch_2d(&map, k2, b2 * AT_T); tma_prefetch_2d(&map, k2, b2 * AT_T + 128);
          tma_prefetch_2d(&map, v2, b2 * AT_T); tma_prefetch_2d(&map, v2, b2 * AT_T + 128);
        }
      }
    }
    mbar_wait(bar_qk, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t idesc = make_idesc(AT_M, AT_T);
      const uint64_t ad = make_desc_sw128(smem_u32(smem + AT_Q_OFF)), bd = make_desc_sw128(smem_u32(smem + AT_K_OFF));
#pragma unroll
      for (int k = 0; k < AT_D / 16; ++k) umma_bf16(tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, k > 0);
      umma_commit(bar_s);
    }
    __syncwarp();
  }

  // ---- softmax: thread = (query row, key half) ----
  const int r = tid & 127;                     // query row = TMEM lane
  const int half = tid >> 7;                   // keys [128 half, +128) / output channels [32 half, +32)
  mbar_wait(bar_s, 0);
  tc_fence_after();
  const uint32_t t_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const int cbase = half * 128;
  float mx = -INFINITY;
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(t_row + (uint32_t)(cbase + c0), v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
  }
  xch[tid] = mx;
  __syncthreads();
  mx = fmaxf(mx, xch[tid ^ 128]);
  const float mxs = mx * p.scale_log2;
  float sum = 0.f;
  uint8_t* prow = smem + AT_P_OFF + r * 128;
  for (int c0 = cbase; c0 < cbase + 128; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(t_row + (uint32_t)c0, v);
    tmem_ld_wait();
    uint8_t* pchunk = prow + (c0 >> 6) * 16384;
    const int c16 = (c0 & 63) >> 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 o4;
      __nv_bfloat162* o2 = (__nv_bfloat162*)&o4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float e0 = ex2_approx(fmaf(__uint_as_float(v[i * 8 + 2 * q]), p.scale_log2, -mxs));
        const float e1 = ex2_approx(fmaf(__uint_as_float(v[i * 8 + 2 * q + 1]), p.scale_log2, -mxs));
        sum += e0 + e1;
        o2[q] = __floats2bfloat162_rn(e0, e1);
      }
      sts_u4(smem_u32(pchunk) + (uint32_t)(((c16 + i) ^ (r & 7)) << 4), o4);     // explicit st.shared (a generic store resolves the space at run time)
    }
  }
  xch[256 + tid] = sum;
  fence_proxy_async();          // P was written through the generic proxy; the MMA reads it through the async proxy
  tc_fence_before();
  __syncthreads();
  sum += xch[256 + (tid ^ 128)];

  if (warp == 0) {
    mbar_wait(bar_v, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t idesc = make_idesc_major(AT_M, AT_D, 0, 1);
      const uint32_t pbase = smem_u32(smem + AT_P_OFF), vbase = smem_u32(smem + AT_V_OFF);
#pragma unroll
      for (int j = 0; j < AT_T / 16; ++j) {
        const uint64_t ad = make_desc_sw128(pbase + (j >> 2) * 16384 + (j & 3) * 32);
        const uint64_t bd = make_desc_sw128_mn(vbase + j * 2048, 1024);
        umma_bf16(tmem, ad, bd, idesc, j > 0);
      }
      umma_commit(bar_o);
    }
    __syncwarp();
  }
  mbar_wait(bar_o, 0);
  tc_fence_after();
  const float inv = 1.0f / sum;
  // each thread stores 32 channels (64 B) of its row as two 256-bit stores (whole sectors)
  {
    const int c0 = half * 32;
    bf16* op = p.out + ((long long)(row0 + mt * AT_M + r)) * p.C + h * AT_D + c0;
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
    stg256(op, o[0], o[1]);
    stg256(op + 16, o[2], o[3]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

#endif  // CFM_TUNING

// ------------------------------------------------------------------------------------------------
// Persistent, pipelined variant (round 2): one CTA per SM walks (sample, head) items; both 128-query tiles of an item
// are in flight together, so K and V are loaded once per head instead of once per query tile, and the next item's
// operands stream in while the current one is in its softmax.
//   warp 0      TMA producer: Q0 | Q1 | K (64 KB) of item j+1 and V (32 KB) of item j into buffers that the PV MMAs of
//               items j-1 / j-2 have released (pv_done barriers) -> ~96 KB in flight per SM throughout
//   warp 1      MMA issuer: S0 = Q0 K^T, S1 = Q1 K^T (128 x 256 x 64 each) into TMEM columns [0, 256) / [256, 512);
//               then O_g = P_g V as the two softmax groups deliver their P
//   warps 2-5   softmax group 0 (query tile 0), one thread per query row: exact two-pass fp32 softmax from TMEM,
//   warps 6-9   softmax group 1 (query tile 1)     bf16 P into swizzled smem, O / rowsum -> bf16 -> global
// Shared memory: three 64 KB buffers rotate through the roles {Q|K of item j, later overlaid by P0 of item j},
// {P1 of item j}, {Q|K of item j+1}: q_j = 2j mod 3 holds Q|K and P0, p_j = q_{j-1} holds P1; V has its own 32 KB.
// 268 M exponentials per block at batch 1024 put the MUFU floor at ~0.06 ms, the DRAM floor (qkv read once, output
// written once) at 0.083 ms; the one-CTA-per-tile kernel above takes 0.197 ms, this one 0.172-0.178 ms (where the rest
// goes: profiles/r02_attn_persist_ncu.txt - waits for S / O, output-store back-pressure, 34 % issue utilisation of the
// one-thread-per-row softmax; a polynomial exp on the FMA pipe and packed f32x2 math measured neutral).
// ------------------------------------------------------------------------------------------------
constexpr int AP_THREADS = 320;
constexpr int AP_BUF = 65536;
constexpr int AP_V_OFF = 3 * AP_BUF;
constexpr int AP_BAR_OFF = AP_V_OFF + 32768;
constexpr int AP_SMEM = AP_BAR_OFF + 256 + 1024;

__global__ void __launch_bounds__(AP_THREADS, 1) attn_tc_persist_kernel(const __grid_constant__ CUtensorMap map, const AttnTcParams p,
                                                                        int n_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_qk = (uint64_t*)(smem + AP_BAR_OFF);   // [2]
  uint64_t* full_v = full_qk + 2;
  uint64_t* bar_s = full_qk + 3;                         // [group][key half]: S_g columns of that half complete
  uint64_t* bar_p = full_qk + 7;                         // [group][key half] 128 arrivals: P_g of that half written
  uint64_t* bar_o = full_qk + 11;                        // [2] O_g complete
  uint64_t* bar_e = full_qk + 13;                        // [2] 128 arrivals: O_g read out of TMEM
  uint64_t* pv_done = full_qk + 15;                      // [2] by item parity: every MMA of the item has retired
  uint32_t* tmem_slot = (uint32_t*)(full_qk + 17);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  pdl_launch_dependents();
  if (tid == 0) {
    prefetch_tmap(&map);
    mbar_init(&full_qk[0], 1); mbar_init(&full_qk[1], 1); mbar_init(full_v, 1);
#pragma unroll
    for (int i = 0; i < 4; ++i) { mbar_init(&bar_s[i], 1); mbar_init(&bar_p[i], 128); }
    mbar_init(&bar_o[0], 1); mbar_init(&bar_o[1], 1); mbar_init(&bar_e[0], 128); mbar_init(&bar_e[1], 128);
    mbar_init(&pv_done[0], 1); mbar_init(&pv_done[1], 1);
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
      int j = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++j) {
        const int b = it / p.heads, h = it - b * p.heads;
        const int qcol = p.new_order ? h * AT_D : h * 3 * AT_D;
        const int kcol = p.new_order ? p.C + h * AT_D : h * 3 * AT_D + AT_D;
        const int vcol = p.new_order ? 2 * p.C + h * AT_D : h * 3 * AT_D + 2 * AT_D;
        const int row0 = b * AT_T;
        uint8_t* qk = smem + ((2 * j) % 3) * AP_BUF;
        if (j >= 2) mbar_wait(&pv_done[j & 1], (uint32_t)(((j - 2) >> 1) & 1));      // this buffer was P1 of item j - 2
        mbar_expect_tx(&full_qk[j & 1], 65536);
        tma_load_2d(qk, &map, &full_qk[j & 1], qcol, row0);
        tma_load_2d(qk + 16384, &map, &full_qk[j & 1], qcol, row0 + 128);
        tma_load_2d(qk + 32768, &map, &full_qk[j & 1], kcol, row0);
        tma_load_2d(qk + 49152, &map, &full_qk[j & 1], kcol, row0 + 128);
        if (j >= 1) mbar_wait(&pv_done[(j - 1) & 1], (uint32_t)(((j - 1) >> 1) & 1));  // V of item j - 1 consumed
        mbar_expect_tx(full_v, 32768);
        tma_load_2d(smem + AP_V_OFF, &map, full_v, vcol, row0);
        tma_load_2d(smem + AP_V_OFF + 16384, &map, full_v, vcol, row0 + 128);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // S_g is issued as two 128-key halves with a barrier each, so the max pass starts a quarter of the way into the S
    // MMAs; O_g = P_g V runs over the first 128 keys while the softmax is still writing the second 128.
    const uint32_t idesc_s = make_idesc(AT_M, AT_T / 2);
    const uint32_t idesc_o = make_idesc_major(AT_M, AT_D, 0, 1);
    const uint32_t vbase = smem_u32(smem + AP_V_OFF);
    int j = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++j) {
      const uint32_t par = (uint32_t)(j & 1);
      const uint32_t qk = smem_u32(smem + ((2 * j) % 3) * AP_BUF);
      const uint32_t p1 = smem_u32(smem + ((2 * j + 1) % 3) * AP_BUF);
      mbar_wait(&full_qk[j & 1], (uint32_t)((j >> 1) & 1));
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        if (j > 0) mbar_wait(&bar_e[m], par ^ 1);        // O_m of the previous item has left TMEM
        tc_fence_after();
        if (elect_one()) {
          const uint64_t qd = make_desc_sw128(qk + m * 16384);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const uint64_t kd = make_desc_sw128(qk + 32768 + hf * 16384);
#pragma unroll
            for (int k = 0; k < AT_D / 16; ++k)
              umma_bf16(tmem + (uint32_t)(m * 256 + hf * 128), qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc_s, k > 0);
            umma_commit(&bar_s[m * 2 + hf]);
          }
        }
        __syncwarp();
      }
#pragma unroll
      for (int step = 0; step < 4; ++step) {
        const int g = step & 1, hf = step >> 1;          // P0 first half, P1 first half, P0 second half, P1 second half
        mbar_wait(&bar_p[g * 2 + hf], par);
        if (step == 0) mbar_wait(full_v, par);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t pbase = g == 0 ? qk : p1;
#pragma unroll
          for (int kk = 0; kk < AT_T / 32; ++kk) {
            const int k = hf * (AT_T / 32) + kk;
            const uint64_t ad = make_desc_sw128(pbase + (k >> 2) * 16384 + (k & 3) * 32);
            const uint64_t bd = make_desc_sw128_mn(vbase + k * 2048, 1024);
            umma_bf16(tmem + (uint32_t)(g * 256), ad, bd, idesc_o, k > 0);
          }
          if (hf == 1) umma_commit(&bar_o[g]);
          if (step == 3) umma_commit(&pv_done[j & 1]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== softmax + epilogue: group g = query tile g, one thread per query row =====================
    const int g = (warp - 2) >> 2;
    const int r = (warp & 3) * 32 + lane;                 // row of the tile = TMEM lane
    const uint32_t t_row = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g * 256);
    int j = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++j) {
      const uint32_t par = (uint32_t)(j & 1);
      const int b = it / p.heads, h = it - b * p.heads;
      uint8_t* prow = smem + ((g == 0 ? 2 * j : 2 * j + 1) % 3) * AP_BUF + r * 128;
      // both passes keep the NEXT 32 scores in flight (tcgen05.ld) while the current 32 are consumed: with one thread per
      // row and two warps per scheduler the TMEM read latency is otherwise the whole cost of the max pass
      float mx = -INFINITY;
      {
        uint32_t va[32], vb[32];
#pragma unroll 1
        for (int c0 = 0; c0 < AT_T; c0 += 64) {
          if ((c0 & 127) == 0) {                          // first touch of a 128-key half: its S MMAs have retired
            mbar_wait(&bar_s[g * 2 + (c0 >> 7)], par);
            tc_fence_after();
            tmem_ld32(t_row + (uint32_t)c0, va);
          }
          tmem_ld_wait();
          tmem_ld32(t_row + (uint32_t)(c0 + 32), vb);
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(va[i]));
          tmem_ld_wait();
          if ((c0 & 127) == 0) tmem_ld32(t_row + (uint32_t)(c0 + 64), va);
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(vb[i]));
        }
      }
      if (g == 0) { mbar_wait(&bar_s[3], par); tc_fence_after(); }   // P0 overlays Q1 | K: S1 must have read them
      const float mxs = mx * p.scale_log2;
      float sum = 0.f;
      auto emit = [&](const uint32_t (&v)[32], int c0) {
        uint8_t* pchunk = prow + (c0 >> 6) * 16384;
        const int c16 = (c0 & 63) >> 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o4;
          __nv_bfloat162* o2 = (__nv_bfloat162*)&o4;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float x0 = fmaf(__uint_as_float(v[i * 8 + 2 * q]), p.scale_log2, -mxs);
            const float x1 = fmaf(__uint_as_float(v[i * 8 + 2 * q + 1]), p.scale_log2, -mxs);
            const float e0 = ex2_approx(x0), e1 = ex2_approx(x1);
            sum += e0 + e1;
            o2[q] = __floats2bfloat162_rn(e0, e1);
          }
          sts_u4(smem_u32(pchunk) + (uint32_t)(((c16 + i) ^ (r & 7)) << 4), o4);     // explicit st.shared (a generic store resolves the space at run time)
        }
      };
      {
        uint32_t va[32], vb[32];
        tmem_ld32(t_row, va);
#pragma unroll 1
        for (int c0 = 0; c0 < AT_T; c0 += 64) {
          tmem_ld_wait();
          tmem_ld32(t_row + (uint32_t)(c0 + 32), vb);
          emit(va, c0);
          tmem_ld_wait();
          if (c0 + 64 < AT_T) tmem_ld32(t_row + (uint32_t)(c0 + 64), va);
          emit(vb, c0 + 32);
          if (c0 == 64) {           // P of the first 128 keys is complete: the PV MMAs over them may start
            fence_proxy_async();    // P went through the generic proxy; the MMA reads it through the async proxy
            tc_fence_before();
            mbar_arrive(&bar_p[g * 2]);
          }
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bar_p[g * 2 + 1]);
      mbar_wait(&bar_o[g], par);
      tc_fence_after();
      const float inv = 1.0f / sum;
      bf16* op = p.out + ((long long)(b * AT_T + g * AT_M + r)) * p.C + h * AT_D;
      uint32_t v0[32], v1[32];
      tmem_ld32(t_row, v0);
      tmem_ld32(t_row + 32u, v1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&bar_e[g]);       // the next item's S_g may overwrite these TMEM columns
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const uint32_t* v = hh ? v1 : v0;
        uint4 o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          __nv_bfloat162* o2 = (__nv_bfloat162*)&o[i];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            o2[q] = __floats2bfloat162_rn(__uint_as_float(v[i * 8 + 2 * q]) * inv, __uint_as_float(v[i * 8 + 2 * q + 1]) * inv);
        }
        stg256(op + hh * 32, o[0], o[1]);
        stg256(op + hh * 32 + 16, o[2], o[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------------------------------------
struct AttnTcPlan { std::map<int, CUtensorMap> maps; };
static std::map<const Op*, AttnTcPlan> g_attn_plans;   // keyed by op address (ops vector is stable after build)
static std::mutex g_attn_mu;   // engines on different host threads share the map (each touches only its own ops' entries)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_attn_encode = nullptr;

bool attn_tc_supported(const Engine& e, const Op& op) {
  if (!e.bf16 || op.kind != OP_ATTN) return false;
  const char* off = tuning_env("CFM_DISABLE_TC_ATTN");
  if (off && off[0] == '1') return false;
  return op.ch == AT_D && op.Hin * op.Win == AT_T;
}

int attn_tc_launch(Engine& e, const Op& op, int B, cudaStream_t st) {
  if (!g_attn_encode) {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { e.err = "cuTensorMapEncodeTiled unavailable"; return CFM_ERR_CUDA; }
    g_attn_encode = (EncodeTiledFn)fn;
  }
  AttnTcPlan* plp;
  { std::lock_guard<std::mutex> lk(g_attn_mu); plp = &g_attn_plans[&op]; }
  AttnTcPlan& pl = *plp;
  const void* qkv = tensor_ptr(e, op.src0, B);
  auto it = pl.maps.find(B);
  if (it == pl.maps.end()) {
    CUtensorMap m;
    const int C3 = 3 * op.Cin;
    cuuint64_t dims[2] = {(cuuint64_t)C3, (cuuint64_t)B * AT_T};
    cuuint64_t strides[1] = {(cuuint64_t)C3 * 2};
    cuuint32_t box[2] = {(cuuint32_t)AT_D, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_attn_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)qkv, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { e.err = "cuTensorMapEncodeTiled(qkv) failed"; return CFM_ERR_CUDA; }
    it = pl.maps.emplace(B, m).first;
  }
  AttnTcParams p{};
  p.heads = op.heads; p.C = op.Cin; p.new_order = e.cfg.use_new_attention_order;
  p.scale_log2 = (1.0f / sqrtf((float)op.ch)) * 1.4426950408889634f;
  p.out = (bf16*)tensor_ptr(e, op.out, B);
#ifdef CFM_TUNING
  if (const char* v = tuning_env("CFM_DISABLE_ATTN_PERSIST"); v && v[0] == '1') {      // round-1 kernel, A/B only
    static DeviceOnce attr;
    if (attr.pending(e.device)) {
      if (cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM) != cudaSuccess) { e.err = "cudaFuncSetAttribute(attn_tc_kernel) failed"; return CFM_ERR_CUDA; }
      attr.done(e.device);
    }
    // two CTAs per SM, two CTAs (query tiles) per (sample, head): one wave covers sm_count pairs
    static const int pf = [] { const char* v2 = tuning_env("CFM_ATTN_PREFETCH"); return v2 ? atoi(v2) : 1; }();
    p.pf_db = pf * e.sm_count / op.heads; p.pf_dh = pf * e.sm_count % op.heads;
    if (B > 65535) { e.err = "attn_tc: batch too large for the grid"; return CFM_ERR_INVALID; }
    LaunchCfg lc(dim3(AT_T / AT_M, op.heads, B), dim3(256), AT_SMEM, st, 1, pdl_enabled());
    if (cudaLaunchKernelEx(&lc.cfg, attn_tc_kernel, it->second, p) != cudaSuccess) { e.err = "attn_tc_kernel launch failed"; return CFM_ERR_CUDA; }
    return 0;
  }
#endif
  static DeviceOnce attr2;
  if (attr2.pending(e.device)) {
    if (cudaFuncSetAttribute(attn_tc_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AP_SMEM) != cudaSuccess) { e.err = "cudaFuncSetAttribute(attn_tc_persist_kernel) failed"; return CFM_ERR_CUDA; }
    attr2.done(e.device);
  }
  const int n_items = B * op.heads;
  LaunchCfg lp(dim3((unsigned)std::min(n_items, e.sm_count)), dim3(AP_THREADS), AP_SMEM, st, 1, pdl_enabled());
  if (cudaLaunchKernelEx(&lp.cfg, attn_tc_persist_kernel, it->second, p, n_items) != cudaSuccess) { e.err = "attn_tc_persist_kernel launch failed"; return CFM_ERR_CUDA; }
  return 0;
}

void attn_tc_release(Engine& e) {
  std::lock_guard<std::mutex> lk(g_attn_mu);
  for (const Op& op : e.ops) {
    auto it = g_attn_plans.find(&op);
    if (it != g_attn_plans.end()) it->second.maps.clear();
  }
}

void attn_tc_forget(Engine& e) {
  std::lock_guard<std::mutex> lk(g_attn_mu);
  for (const Op& op : e.ops) g_attn_plans.erase(&op);
}

}  // namespace cfm
