// Memory-bound bf16 kernels of the tensor-core path: GroupNorm(+FiLM)(+SiLU), nearest/avg resample,
// the 3-channel edge convolutions.  All are HBM-bound: 16-byte vector accesses, each tensor read once.
#include <algorithm>
#include "engine.h"
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cfm {

void* tensor_ptr(const Engine& e, int id, int B);

// ------------------------------------------------------------------------------------------------
// GroupNorm32 over NHWC bf16, staged through shared memory.
// An "item" is one (sample, channel slab); a slab is a whole number of groups and of 8-channel
// 16-byte vectors.  One CTA (or a cluster of `cs` CTAs that split the item's pixels) owns an item:
//   1. every thread copies its vectors global -> shared with cp.async (no register staging, so a
//      CTA costs ~40 registers/thread and several CTAs per SM keep their loads in flight while others
//      reduce or store - the kernel is HBM-bound and bytes in flight are what matters);
//   2. each thread reduces the vectors it copied itself (no barrier needed to read them back); the
//      per-channel sums are combined in a fixed order - shuffles + one smem pass, DSMEM across the
//      cluster - so the result is bit-reproducible (no atomics);
//   3. the same vectors are read back from shared memory, normalised / FiLM-modulated / SiLU-activated
//      and stored with 16-byte coalesced stores: one global read, one global write.
// ------------------------------------------------------------------------------------------------
struct GnFastArgs {
  const bf16* src0; const bf16* src1; int C0, C1;
  int HW, cpg, slab;                 // slab channels per item (multiple of 8 and of cpg)
  int n_items;
  int tpi, ipc;                      // threads per item, items per CTA (ipc > 1 only without a cluster)
  int cs;                            // CTAs (of one cluster) sharing an item's pixels; 1 = no cluster
  int shfl;                          // 1: vectors-per-pixel is a power of two <= 32 -> warp-shuffle pre-reduction
  int data_bytes;                    // staged bytes per CTA
  const float* gamma; const float* beta; float eps; int silu;
  const float* film; int film_stride; const int* film_row;
  bf16* out;
};

constexpr int GN_MAX_SLAB = 128;
constexpr int GN_MAX_THREADS = 256;

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint4 lds_u4(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}

__global__ void __launch_bounds__(GN_MAX_THREADS) groupnorm_bf16_kernel(GnFastArgs a) {
  extern __shared__ __align__(16) uint8_t gn_smem[];
  // layout: data[ipc][data_bytes] | red[rows][17] | ch_sum[ipc][128] | ch_sq | ch_scale | ch_shift
  // Small feature maps pack `ipc` items into one CTA (tpi threads each) so a CTA still moves >= 16 KB.
  const int T = blockDim.x;
  const int vpp = a.slab >> 3;
  const int tpi = a.tpi, ipc = a.ipc;
  const int wpi = tpi >> 5;                                // warps per item
  const int red_rows = a.shfl ? (T >> 5) * vpp : T;
  float* red = (float*)(gn_smem + ipc * a.data_bytes);     // rows of 16 partials, stride 17 (bank-conflict-free per-thread rows)
  float* ch_sum = red + red_rows * 17;
  float* ch_sq = ch_sum + ipc * GN_MAX_SLAB;
  float* ch_scale = ch_sq + ipc * GN_MAX_SLAB;
  float* ch_shift = ch_scale + ipc * GN_MAX_SLAB;
  // cluster mode: recv[rank][channel] = {sum, sumsq} pushed by every CTA of the cluster, counted by xbar
  float* recv = ch_shift + ipc * GN_MAX_SLAB;
  uint64_t* xbar = (uint64_t*)(((uintptr_t)(recv + a.cs * a.slab * 2) + 7) & ~(uintptr_t)7);

  pdl_launch_dependents();
  if (a.cs > 1) {
    if (threadIdx.x == 0) {
      mbar_init(xbar, 1);
      fence_barrier_init();
      mbar_expect_tx(xbar, (uint32_t)(a.cs * a.slab * 8));
    }
    cluster_sync_relaxed();          // every CTA's barrier exists before anybody pushes to it
  }
  pdl_wait();
  const int C = a.C0 + a.C1;
  const int slabs = C / a.slab;
  const int il = threadIdx.x / tpi;                        // item within the CTA
  const int t = threadIdx.x - il * tpi;                    // thread within the item
  const int crank = a.cs > 1 ? (int)cluster_ctarank() : 0;
  const int item = a.cs > 1 ? (int)(blockIdx.x / a.cs) : (int)blockIdx.x * ipc + il;
  const bool active = item < a.n_items;
  const int HWl = a.HW / a.cs;                           // pixels owned by this CTA
  const int b = active ? item / slabs : 0, sl = active ? item - b * slabs : 0;
  const int c_base = sl * a.slab;
  const int q = t % vpp;                                 // tpi % vpp == 0 -> fixed vector slot (8 channels) per thread
  const int cq = c_base + q * 8;
  const bf16* sp; int sC, sc;
  if (cq < a.C0) { sp = a.src0; sC = a.C0; sc = cq; } else { sp = a.src1; sC = a.C1; sc = cq - a.C0; }
  const long long pix0 = (long long)b * a.HW + (long long)crank * HWl;
  const int p0 = t / vpp, pstep = tpi / vpp;             // this thread's pixels: p0 + k * pstep
  const int nvec = active ? HWl * vpp : 0;
  const uint32_t data_addr = smem_u32(gn_smem) + (uint32_t)(il * a.data_bytes);

  {
    const bf16* lp = sp + (pix0 + p0) * sC + sc;
    const long long lstep = (long long)pstep * sC;
    for (int v = t; v < nvec; v += tpi, lp += lstep) cp_async16(data_addr + (uint32_t)v * 16u, lp);
  }
  cp_async_wait_all();
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; ss[j] = 0.f; }
#pragma unroll 4
  for (int v = t; v < nvec; v += tpi) {
    const uint4 r = lds_u4(data_addr + (uint32_t)v * 16u);
    const __nv_bfloat162* h2 = (const __nv_bfloat162*)&r;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(h2[j]);
      s[2 * j] += f.x; ss[2 * j] = fmaf(f.x, f.x, ss[2 * j]);
      s[2 * j + 1] += f.y; ss[2 * j + 1] = fmaf(f.y, f.y, ss[2 * j + 1]);
    }
  }
  if (a.shfl) {
    // lanes that share a vector slot are vpp apart: butterfly over the lane bits above log2(vpp)
    for (int m = vpp; m < 32; m <<= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += __shfl_xor_sync(0xffffffffu, s[j], m); ss[j] += __shfl_xor_sync(0xffffffffu, ss[j], m); }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane < vpp) {
      float* rp = red + (w * vpp + lane) * 17;
#pragma unroll
      for (int j = 0; j < 8; ++j) { rp[j] = s[j]; rp[8 + j] = ss[j]; }
    }
  } else {
    float* rp = red + threadIdx.x * 17;
#pragma unroll
    for (int j = 0; j < 8; ++j) { rp[j] = s[j]; rp[8 + j] = ss[j]; }
  }
  __syncthreads();
  for (int c = t; c < a.slab; c += tpi) {
    // fixed-order sum of the partials that belong to channel c of this item's slab
    const int qq = c >> 3, j = c & 7;
    float ts = 0.f, tq = 0.f;
    if (a.shfl) { for (int w = il * wpi; w < (il + 1) * wpi; ++w) { const float* rp = red + (w * vpp + qq) * 17; ts += rp[j]; tq += rp[8 + j]; } }
    else { for (int u = il * tpi + qq; u < (il + 1) * tpi; u += vpp) { ts += red[u * 17 + j]; tq += red[u * 17 + 8 + j]; } }
    ch_sum[il * GN_MAX_SLAB + c] = ts; ch_sq[il * GN_MAX_SLAB + c] = tq;
  }
  __syncthreads();
  if (a.cs > 1) {
    // combine the per-CTA channel sums across the cluster: every CTA PUSHES its partials into every CTA's recv[rank]
    // (st.async, counted by the receiver's mbarrier) and sums what it received in rank order - no cluster barrier and
    // no release fence in the item's critical path (the two barrier.cluster rounds this replaces cost ~17 % of the
    // kernel's stall samples: membar 11 %, barrier wait 6 %).  Cluster CTAs hold one item and have >= slab threads.
    if (t < a.slab) {
      const float ts = ch_sum[t], tq = ch_sq[t];
      const uint32_t dst = smem_u32(&recv[(crank * a.slab + t) * 2]), bar = smem_u32(xbar);
      for (int r = 0; r < a.cs; ++r) st_async_v2f32(mapa_u32(dst, r), ts, tq, mapa_u32(bar, r));
    }
    mbar_wait(xbar, 0);
    if (t < a.slab) {
      float ts = 0.f, tq = 0.f;
      for (int r = 0; r < a.cs; ++r) { ts += recv[(r * a.slab + t) * 2]; tq += recv[(r * a.slab + t) * 2 + 1]; }
      ch_sum[t] = ts; ch_sq[t] = tq;
    }
    __syncthreads();
  }
  for (int c = t; c < a.slab; c += tpi) {
    const int g0 = (c / a.cpg) * a.cpg;
    float gs = 0.f, gq = 0.f;
    for (int j = 0; j < a.cpg; ++j) { gs += ch_sum[il * GN_MAX_SLAB + g0 + j]; gq += ch_sq[il * GN_MAX_SLAB + g0 + j]; }
    const float inv_n = 1.0f / (float)(a.cpg * a.HW);
    const float mean = gs * inv_n;
    const float var = fmaxf(gq * inv_n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + a.eps);
    float sc_ = rstd * a.gamma[c_base + c];
    float sh_ = a.beta[c_base + c] - mean * sc_;
    if (a.film) {
      const float* f = a.film + (long long)a.film_row[b] * a.film_stride;
      const float m = 1.0f + f[c_base + c];
      sc_ *= m; sh_ = sh_ * m + f[C + c_base + c];
    }
    if (a.silu) { sc_ *= 0.5f; sh_ *= 0.5f; }          // the activation works on h = y/2: silu(y) = h*tanh(h) + h
    ch_scale[il * GN_MAX_SLAB + c] = sc_; ch_shift[il * GN_MAX_SLAB + c] = sh_;
  }
  __syncthreads();
  float sc8[8], sh8[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc8[j] = ch_scale[il * GN_MAX_SLAB + q * 8 + j]; sh8[j] = ch_shift[il * GN_MAX_SLAB + q * 8 + j]; }
  bf16* op = a.out + (pix0 + p0) * C + cq;
  const long long ostep = (long long)pstep * C;
#pragma unroll 4
  for (int v = t; v < nvec; v += tpi, op += ostep) {
    const uint4 r = lds_u4(data_addr + (uint32_t)v * 16u);
    const __nv_bfloat162* h2 = (const __nv_bfloat162*)&r;
    uint4 o4;
    __nv_bfloat162* o2 = (__nv_bfloat162*)&o4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(h2[j]);
      float y0 = fmaf(f.x, sc8[2 * j], sh8[2 * j]), y1 = fmaf(f.y, sc8[2 * j + 1], sh8[2 * j + 1]);
      if (a.silu) { y0 = silu_from_half(y0); y1 = silu_from_half(y1); }
      o2[j] = __floats2bfloat162_rn(y0, y1);
    }
    *(uint4*)op = o4;
  }
}

static int gcd_i(int a, int b) { return b ? gcd_i(b, a % b) : a; }

struct GnGeom { int cpg, slab, threads, tpi, ipc, cs, shfl, data_bytes; size_t smem; bool ok; };

static GnGeom gn_geometry(const Op& op) {
  GnGeom g{};
  const int C = op.Cin, HW = op.Hin * op.Win;
  g.cpg = C / 32;
  const int base = g.cpg / gcd_i(g.cpg, 8) * 8;        // lcm(cpg, 8): smallest legal slab
  if (base > GN_MAX_SLAB || C % base) return g;
  static const int target = [] { const char* v = tuning_env("CFM_GN_ITEM_BYTES"); return v ? atoi(v) : 96 * 1024; }();
  // slab: a multiple of `base` dividing C, preferably a whole number of 64-byte DRAM bursts per pixel (32 channels;
  // e.g. 96 for C = 384, where 48-channel slabs would straddle bursts) and at most 64 channels when that works
  static const int pref_slab = [] { const char* v = tuning_env("CFM_GN_PREF_SLAB"); return v ? atoi(v) : 64; }();
  // ... but 128 channels (256-byte rows) when such an item still fits one CTA without a cluster (16x16 maps: -4 %)
  const int pref = (long long)HW * GN_MAX_SLAB * 2 <= target ? std::max(pref_slab, GN_MAX_SLAB) : pref_slab;
  int slab = 0;
  for (int sl = base; sl <= GN_MAX_SLAB; sl += base)
    if (C % sl == 0 && sl % 32 == 0) { if (slab == 0 || sl <= pref) slab = sl; }
  if (slab == 0) { slab = base; while (slab * 2 <= 64 && C % (slab * 2) == 0) slab *= 2; }
  // bring the per-CTA bytes to the target: split the pixels over a cluster first, then narrow the slab (>= 64 B per pixel)
  int cs = 1;
  while ((long long)HW * slab * 2 / cs > target && cs < 8 && HW % (cs * 2) == 0 && HW / (cs * 2) >= 8) cs *= 2;
  while ((long long)HW * slab * 2 / cs > target && slab % 64 == 0 && (slab / 2) % base == 0) slab /= 2;
  const long long bytes = (long long)HW * slab * 2 / cs;
  if (bytes > 160 * 1024) return g;
  const int vpp = slab / 8;
  const int unit = 32 / gcd_i(32, vpp) * vpp;          // lcm(32, vpp): whole warps, multiple of vpp
  if (unit > GN_MAX_THREADS) return g;                 // e.g. 18 channels per group (C = 576): 9 vectors per pixel -> generic kernel
  const int nvec = (int)(bytes / 16);
  int threads = ((nvec + 7) / 8 + unit - 1) / unit * unit;   // ~8 vectors per thread
  threads = std::min(std::max(threads, cs > 1 ? std::max(unit, (GN_MAX_SLAB + unit - 1) / unit * unit) : unit), GN_MAX_THREADS / unit * unit);
  // small maps: several items per CTA so that a CTA still moves ~16 KB
  int ipc = 1;
  if (cs == 1) ipc = (int)std::max<long long>(1, std::min<long long>(GN_MAX_THREADS / threads, 16384 / std::max<long long>(bytes, 1)));
  g.slab = slab; g.cs = cs; g.tpi = threads; g.ipc = ipc; g.threads = threads * ipc; g.data_bytes = (int)bytes;
  g.shfl = (vpp & (vpp - 1)) == 0 && vpp <= 32;
  const int red_rows = g.shfl ? (g.threads / 32) * vpp : g.threads;
  g.smem = (size_t)bytes * ipc + sizeof(float) * ((size_t)red_rows * 17 + 4 * (size_t)ipc * GN_MAX_SLAB);
  if (cs > 1) g.smem += sizeof(float) * (size_t)cs * slab * 2 + 32;    // pushed partials of every rank + the mbarrier that counts them
  g.ok = true;
  return g;
}

bool gn_bf16_supported(const Engine& e, const Op& op) {
  if (!e.bf16 || op.kind != OP_GN) return false;
  const char* off = tuning_env("CFM_DISABLE_FAST_GN");
  if (off && off[0] == '1') return false;
  const GnGeom g = gn_geometry(op);
  if (!g.ok || g.smem > 200 * 1024) return false;
  if (e.tensors[op.src0].C % 8) return false;
  return true;
}

int gn_bf16_launch(Engine& e, const Op& op, int B, cudaStream_t st) {
  static DeviceOnce attr;
  if (attr.pending(e.device)) {
    if (cudaFuncSetAttribute(groupnorm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
      e.err = "cudaFuncSetAttribute(groupnorm_bf16_kernel) failed"; return CFM_ERR_CUDA;
    }
    attr.done(e.device);
  }
  const GnGeom g = gn_geometry(op);
  GnFastArgs a{};
  a.cpg = g.cpg; a.slab = g.slab; a.cs = g.cs; a.shfl = g.shfl; a.data_bytes = g.data_bytes; a.tpi = g.tpi; a.ipc = g.ipc;
  a.src0 = (const bf16*)tensor_ptr(e, op.src0, B); a.C0 = e.tensors[op.src0].C;
  a.src1 = (const bf16*)tensor_ptr(e, op.src1, B); a.C1 = op.src1 >= 0 ? e.tensors[op.src1].C : 0;
  a.HW = op.Hin * op.Win;
  a.n_items = B * (op.Cin / g.slab);
  a.gamma = op.gamma; a.beta = op.beta; a.eps = 1e-5f; a.silu = op.silu;
  if (op.film) { a.film = e.emb_out + op.emb_off; a.film_stride = e.emb_total; a.film_row = e.row_of_sample; }
  a.out = (bf16*)tensor_ptr(e, op.out, B);
  LaunchCfg lc(dim3((unsigned)(g.cs > 1 ? a.n_items * g.cs : (a.n_items + g.ipc - 1) / g.ipc)), dim3(g.threads), g.smem, st, g.cs, pdl_enabled());
  if (cudaLaunchKernelEx(&lc.cfg, groupnorm_bf16_kernel, a) != cudaSuccess) { e.err = "groupnorm launch failed"; return CFM_ERR_CUDA; }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// nearest x2 upsample / 2x2 average pool, 8 channels (16 B) per thread
// ------------------------------------------------------------------------------------------------
__global__ void resample_bf16_kernel(const bf16* __restrict__ src, bf16* __restrict__ out, int B, int Hin, int Win, int C, int up) {
  const int Ho = up ? Hin * 2 : Hin / 2, Wo = up ? Win * 2 : Win / 2;
  const int cv = C / 8;
  const long long total = (long long)B * Ho * Wo * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * 8;
    long long p = i / cv;
    const int ox = (int)(p % Wo); p /= Wo;
    const int oy = (int)(p % Ho);
    const int b = (int)(p / Ho);
    uint4 o4;
    if (up) {
      o4 = __ldg((const uint4*)(src + (((long long)b * Hin + (oy >> 1)) * Win + (ox >> 1)) * C + c));
    } else {
      const bf16* base = src + (((long long)b * Hin + oy * 2) * Win + ox * 2) * C + c;
      const uint4 r0 = __ldg((const uint4*)base), r1 = __ldg((const uint4*)(base + C));
      const uint4 r2 = __ldg((const uint4*)(base + (long long)Win * C)), r3 = __ldg((const uint4*)(base + (long long)Win * C + C));
      const __nv_bfloat162 *a0 = (const __nv_bfloat162*)&r0, *a1 = (const __nv_bfloat162*)&r1, *a2 = (const __nv_bfloat162*)&r2, *a3 = (const __nv_bfloat162*)&r3;
      __nv_bfloat162* o2 = (__nv_bfloat162*)&o4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f0 = __bfloat1622float2(a0[j]), f1 = __bfloat1622float2(a1[j]), f2 = __bfloat1622float2(a2[j]), f3 = __bfloat1622float2(a3[j]);
        o2[j] = __floats2bfloat162_rn(0.25f * (f0.x + f1.x + f2.x + f3.x), 0.25f * (f0.y + f1.y + f2.y + f3.y));
      }
    }
    *(uint4*)(out + i * 8) = o4;
  }
}

bool resample_bf16_supported(const Engine& e, const Op& op) { return e.bf16 && op.kind == OP_RESAMPLE && op.Cin % 8 == 0; }

int resample_bf16_launch(Engine& e, const Op& op, int B, cudaStream_t st) {
  const long long total = (long long)B * e.tensors[op.out].elems() / 8;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)e.sm_count * 32);
  resample_bf16_kernel<<<blocks, 256, 0, st>>>((const bf16*)tensor_ptr(e, op.src0, B), (bf16*)tensor_ptr(e, op.out, B), B, op.Hin, op.Win, op.Cin, op.up);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Network head: 3x3 conv, bf16 NHWC in (C % 8 == 0), few output channels (<= 4), fp32 NCHW out.
// One thread per output pixel; weights (fp32) broadcast from shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int HEAD_MAX_COUT = 4;

__global__ void __launch_bounds__(128) head_conv_kernel(const bf16* __restrict__ src, const float* __restrict__ w /*[9*C][Cout]*/,
                                                        const float* __restrict__ bias, float* __restrict__ out,
                                                        int B, int H, int W, int C, int Cout) {
  extern __shared__ float ws[];      // [9*C][HEAD_MAX_COUT]
  for (int i = threadIdx.x; i < 9 * C * HEAD_MAX_COUT; i += blockDim.x) {
    const int k = i / HEAD_MAX_COUT, o = i % HEAD_MAX_COUT;
    ws[i] = o < Cout ? w[(long long)k * Cout + o] : 0.f;
  }
  __syncthreads();
  const long long total = (long long)B * H * W;
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= total) return;
  const int x = (int)(m % W), y = (int)((m / W) % H), b = (int)(m / ((long long)W * H));
  float acc[HEAD_MAX_COUT];
#pragma unroll
  for (int o = 0; o < HEAD_MAX_COUT; ++o) acc[o] = 0.f;
  for (int tap = 0; tap < 9; ++tap) {
    const int iy = y + tap / 3 - 1, ix = x + tap % 3 - 1;
    if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
    const bf16* sp = src + (((long long)b * H + iy) * W + ix) * C;
    const float4* wp = (const float4*)(ws + (long long)tap * C * HEAD_MAX_COUT);
#pragma unroll 4
    for (int c = 0; c < C; c += 8) {
      const uint4 raw = __ldg((const uint4*)(sp + c));
      const __nv_bfloat162* h2 = (const __nv_bfloat162*)&raw;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h2[j]);
        const float4 w0 = wp[c + 2 * j], w1 = wp[c + 2 * j + 1];
        acc[0] = fmaf(f.x, w0.x, acc[0]); acc[1] = fmaf(f.x, w0.y, acc[1]); acc[2] = fmaf(f.x, w0.z, acc[2]); acc[3] = fmaf(f.x, w0.w, acc[3]);
        acc[0] = fmaf(f.y, w1.x, acc[0]); acc[1] = fmaf(f.y, w1.y, acc[1]); acc[2] = fmaf(f.y, w1.z, acc[2]); acc[3] = fmaf(f.y, w1.w, acc[3]);
      }
    }
  }
  const long long hw = (long long)H * W;
  for (int o = 0; o < Cout; ++o) out[((long long)b * Cout + o) * hw + (long long)y * W + x] = acc[o] + bias[o];
}

bool head_conv_supported(const Engine& e, const Op& op) {
  return e.bf16 && op.kind == OP_CONV && op.out_is_output && !op.src_is_input && op.ks == 3 && op.stride == 1 && !op.ups &&
         op.src1 < 0 && op.skip0 < 0 && op.res0 < 0 && op.emb_off < 0 && op.Cout <= HEAD_MAX_COUT && op.Cin % 8 == 0 &&
         (size_t)9 * op.Cin * HEAD_MAX_COUT * 4 <= 160 * 1024;
}

int head_conv_launch(Engine& e, const Op& op, int B, float* out, cudaStream_t st) {
  const size_t smem = (size_t)9 * op.Cin * HEAD_MAX_COUT * 4;
  static DeviceOnce attr;
  if (attr.pending(e.device)) {
    if (cudaFuncSetAttribute(head_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess) { e.err = "cudaFuncSetAttribute(head_conv_kernel) failed"; return CFM_ERR_CUDA; }
    attr.done(e.device);
  }
  const long long total = (long long)B * op.Hout * op.Wout;
  head_conv_kernel<<<(unsigned)((total + 127) / 128), 128, smem, st>>>((const bf16*)tensor_ptr(e, op.src0, B), op.w_main, op.bias, out, B, op.Hout, op.Wout, op.Cin, op.Cout);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Network stem: 3x3 conv from the fp32 NCHW input (x [+ cond], few channels) to bf16 NHWC.
// One thread per (pixel, 8 output channels); the <= 9*8 input taps are loaded once per thread.
// ------------------------------------------------------------------------------------------------
constexpr int STEM_MAX_CIN = 8;

// One thread per (pixel, 8 output channels); the Cout/8 threads of a pixel are adjacent lanes, so their
// input-tap loads are one broadcast transaction and their 16-byte stores tile the pixel's NHWC row
// contiguously (fully coalesced writes - this kernel is bound by writing the [B,H,W,Cout] tensor).
__global__ void __launch_bounds__(256) stem_conv_kernel(const float* __restrict__ x0, const float* __restrict__ x1, int C0, int C1,
                                                        const float* __restrict__ w /*[9*Cin][Cout]*/, const float* __restrict__ bias,
                                                        bf16* __restrict__ out, int B, int H, int W, int Cout) {
  extern __shared__ float sw[];        // [9*Cin][Cout] then bias[Cout]
  const int Cin = C0 + C1;
  const int K = 9 * Cin;
  for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) sw[i] = w[i];
  float* sb = sw + K * Cout;
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sb[i] = bias[i];
  __syncthreads();
  const int cv = Cout >> 3;
  const long long total = (long long)B * H * W * cv;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int co = (int)(i % cv) << 3;
    const long long m = i / cv;
    const int x = (int)(m % W), y = (int)((m / W) % H), b = (int)(m / ((long long)W * H));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = sb[co + j];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int iy = y + tap / 3 - 1, ix = x + tap % 3 - 1;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
      for (int c = 0; c < Cin; ++c) {
        const float v = (c < C0) ? __ldg(x0 + (((long long)b * C0 + c) * H + iy) * W + ix)
                                 : __ldg(x1 + (((long long)b * C1 + (c - C0)) * H + iy) * W + ix);
        const float4 w0 = *(const float4*)(sw + (tap * Cin + c) * Cout + co);
        const float4 w1 = *(const float4*)(sw + (tap * Cin + c) * Cout + co + 4);
        acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
        acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]); acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
      }
    }
    uint4 o4;
    __nv_bfloat162* o2 = (__nv_bfloat162*)&o4;
#pragma unroll
    for (int j = 0; j < 4; ++j) o2[j] = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
    *(uint4*)(out + m * Cout + co) = o4;
  }
}

bool stem_conv_supported(const Engine& e, const Op& op) {
  return e.bf16 && op.kind == OP_CONV && op.src_is_input && !op.out_is_output && op.ks == 3 && op.stride == 1 && !op.ups &&
         op.skip0 < 0 && op.res0 < 0 && op.emb_off < 0 && op.Cin <= STEM_MAX_CIN && op.Cout % 32 == 0 &&
         (size_t)(9 * op.Cin + 1) * op.Cout * 4 <= 96 * 1024;
}

int stem_conv_launch(Engine& e, const Op& op, int B, const float* x, const float* cond, cudaStream_t st) {
  const int cx = cond ? e.x_channels() : e.cfg.in_channels;
  const size_t smem = (size_t)(9 * op.Cin + 1) * op.Cout * 4;
  static DeviceOnce attr;
  if (attr.pending(e.device)) {
    if (cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess) { e.err = "cudaFuncSetAttribute(stem_conv_kernel) failed"; return CFM_ERR_CUDA; }
    attr.done(e.device);
  }
  const long long total = (long long)B * op.Hout * op.Wout * (op.Cout / 8);
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)e.sm_count * 8);
  stem_conv_kernel<<<blocks, 256, smem, st>>>(x, cond, cx, e.cfg.in_channels - cx, op.w_main, op.bias,
                                                                     (bf16*)tensor_ptr(e, op.out, B), B, op.Hout, op.Wout, op.Cout);
  return 0;
}


// ------------------------------------------------------------------------------------------------
// Stem im2col: fp32 NCHW network input (x [+ cond]) -> bf16 NHWC rows of K_pad values per pixel:
//   [0, 9*Cin)        hi = bf16(x_tap)            (tap-major, channel-minor; zero outside the image)
//   [9*Cin, 18*Cin)   lo = bf16(x_tap - hi)       (second bf16 term: the input keeps ~16 mantissa bits)
//   rest              0
// so the 3x3 stem conv becomes a K_pad-deep 1x1 GEMM on the tensor cores (weights duplicated for hi/lo).
// ------------------------------------------------------------------------------------------------
// One CTA per strip of R image rows (R = the largest of 16, 8, 4 whose patch fits 48 KB; measured 0.092 / 0.083 / 0.080 / 0.085 ms at 4 / 8 / 16 / 32 rows): the (rows + 2) x (W + 2) x Cin input patch is staged in shared
// memory with coalesced NCHW reads (zero halo), then every thread assembles 16-byte vectors of 8 consecutive K
// values - consecutive threads write consecutive vectors, so the [B,H,W,K_pad] tensor (what bounds this kernel)
// is written in full 128-byte lines.
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ x0, const float* __restrict__ x1, int C0, int C1,
                                                          bf16* __restrict__ out, int B, int H, int W, int Kpad, int IM2COL_ROWS) {
  extern __shared__ float im_smem[];     // patch[Cin][rows + 2][W + 2] | koff[Kpad] (int)
  const int Cin = C0 + C1, K9 = 9 * Cin;
  const int strips = (H + IM2COL_ROWS - 1) / IM2COL_ROWS;
  const int b = blockIdx.x / strips, y0 = (blockIdx.x % strips) * IM2COL_ROWS;
  const int rows = min(IM2COL_ROWS, H - y0);
  const int PW = W + 2, PH = IM2COL_ROWS + 2;
  float* patch = im_smem;
  int* koff = (int*)(im_smem + Cin * PH * PW);
  for (int i = threadIdx.x; i < Cin * PH * PW; i += blockDim.x) {
    const int px = i % PW, py = (i / PW) % PH, c = i / (PW * PH);
    const int ix = px - 1, iy = y0 + py - 1;
    float v = 0.f;
    if (ix >= 0 && ix < W && iy >= 0 && iy < H && py < rows + 2)
      v = (c < C0) ? __ldg(x0 + (((long long)b * C0 + c) * H + iy) * W + ix)
                   : __ldg(x1 + (((long long)b * C1 + (c - C0)) * H + iy) * W + ix);
    patch[i] = v;
  }
  // K index -> offset of the tap inside the patch relative to the pixel's top-left halo corner; bit 30 = "lo" term, -1 = zero
  for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
    int kk = k, lo = 0;
    if (kk >= K9) { kk -= K9; lo = 1; }
    int off = -1;
    if (kk < K9) { const int tap = kk / Cin, c = kk - tap * Cin; off = (c * PH + tap / 3) * PW + tap % 3; if (lo) off |= 1 << 30; }
    koff[k] = off;
  }
  __syncthreads();
  const int cv = Kpad >> 3;
  const int total = rows * W * cv;
  bf16* obase = out + (((long long)b * H + y0) * W) * Kpad;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int k0 = (i % cv) << 3;
    const int m = i / cv;                  // pixel within the strip
    const int x = m % W, ry = m / W;
    const int pbase = ry * PW + x;
    uint4 o4;
    bf16* ob = (bf16*)&o4;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int off = koff[k0 + j];
      float v = 0.f;
      if (off >= 0) {
        v = patch[pbase + (off & 0x3fffffff)];
        if (off >> 30) v = v - __bfloat162float(__float2bfloat16_rn(v));
      }
      ob[j] = __float2bfloat16_rn(v);
    }
    *(uint4*)(obase + (long long)i * 8) = o4;
  }
}

int stem_im2col_launch(Engine& e, const Op& op, int B, const float* x, const float* cond, cudaStream_t st) {
  const int cx = cond ? e.x_channels() : e.cfg.in_channels;
  auto smem_for = [&](int r) { return sizeof(float) * ((size_t)op.Cin * (r + 2) * (op.Win + 2) + op.Cout); };
  static const int max_rows = [] { const char* v = tuning_env("CFM_IM2COL_ROWS"); return v ? atoi(v) : 16; }();
  int R = 4;
  for (int r = 32; r >= 4; r >>= 1)
    if (r <= max_rows && r <= std::max(4, op.Hin) && smem_for(r) <= 48 * 1024) { R = r; break; }
  const int strips = (op.Hin + R - 1) / R;
  const size_t smem = smem_for(R);
  if (smem > 48 * 1024) { e.err = "stem im2col patch does not fit shared memory"; return CFM_ERR_INVALID; }
  stem_im2col_kernel<<<B * strips, 256, smem, st>>>(x, cond, cx, e.cfg.in_channels - cx, (bf16*)tensor_ptr(e, op.out, B), B, op.Hin, op.Win, op.Cout, R);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Network head, second half.  The 3x3 head conv (C -> <= 3 channels) runs as a 1x1 tensor-core GEMM that produces, per
// pixel, the 9 x Cout per-tap partial products Y[p][tap * Cout + o] = sum_c a[p][c] w[o][c][tap] (fp32, 32 values per
// pixel); this kernel adds the bias and sums the 9 taps of the 3x3 neighbourhood (zero outside the image) into the
// fp32 NCHW network output.  The activation tensor is read once instead of nine times.
// ------------------------------------------------------------------------------------------------
// One CTA per strip of image rows (<= 256 pixels): the strip's partial products plus one halo row above and below
// are staged in shared memory with coalesced 16-byte loads (28 of the 32 floats of a pixel; row stride 29 floats keeps
// the per-pixel reads of a warp on distinct banks), then each thread sums its pixel's 9 taps.
constexpr int HG_STRIDE = 29;

__global__ void __launch_bounds__(256) head_gather_kernel(const float* __restrict__ Y, const float* __restrict__ bias,
                                                          float* __restrict__ out, int B, int H, int W, int Cout, int R) {
  extern __shared__ float hg_smem[];       // [(R + 2) * W][HG_STRIDE]
  const int strips = (H + R - 1) / R;
  const int b = blockIdx.x / strips, y0 = (blockIdx.x % strips) * R;
  const int rows = min(R, H - y0);
  const int npix = (rows + 2) * W;          // halo rows y0 - 1 .. y0 + rows
  const int n4 = (9 * Cout + 3) >> 2;       // 16-byte vectors of a pixel that carry partial products (7 for 3 channels)
  for (int i = threadIdx.x; i < npix * n4; i += blockDim.x) {
    const int px = i / n4, k4 = i - px * n4;
    const int iy = y0 - 1 + px / W, ix = px % W;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (iy >= 0 && iy < H) v = __ldg((const float4*)(Y + ((((long long)b * H + iy) * W + ix) << 5)) + k4);
    float* sp = hg_smem + px * HG_STRIDE + 4 * k4;
    sp[0] = v.x; sp[1] = v.y; sp[2] = v.z; sp[3] = v.w;
  }
  __syncthreads();
  const long long hw = (long long)H * W;
  for (int m = threadIdx.x; m < rows * W; m += blockDim.x) {
    const int ry = m / W, x = m - ry * W;
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int ix = x + tap % 3 - 1;
      if (ix < 0 || ix >= W) continue;                       // rows outside the image were staged as zeros
      const float* sp = hg_smem + ((ry + tap / 3) * W + ix) * HG_STRIDE + tap * Cout;
      for (int o = 0; o < Cout; ++o) acc[o] += sp[o];
    }
    for (int o = 0; o < Cout; ++o) out[((long long)b * Cout + o) * hw + (long long)(y0 + ry) * W + x] = acc[o] + bias[o];
  }
}

int head_gather_launch(Engine& e, const Op& op, int B, float* out, cudaStream_t st) {
  const int R = std::max(1, std::min(op.Hin, 256 / op.Win));
  const int strips = (op.Hin + R - 1) / R;
  const size_t smem = sizeof(float) * (size_t)(R + 2) * op.Win * HG_STRIDE;
  static DeviceOnce attr;
  if (attr.pending(e.device)) {
    if (cudaFuncSetAttribute(head_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess) { e.err = "cudaFuncSetAttribute(head_gather_kernel) failed"; return CFM_ERR_CUDA; }
    attr.done(e.device);
  }
  if (smem > 96 * 1024) { e.err = "head gather strip does not fit shared memory"; return CFM_ERR_INVALID; }
  head_gather_kernel<<<B * strips, 256, smem, st>>>((const float*)tensor_ptr(e, op.src0, B), op.bias, out, B, op.Hin, op.Win, op.Cout, R);
  return 0;
}

}  // namespace cfm
