// Memory-bound bf16 kernels of the tensor-core path: GroupNorm(+FiLM)(+SiLU), nearest/avg resample,
// the 3-channel edge convolutions.  All are HBM-bound: 16-byte vector accesses, each tensor read once.
#include <algorithm>
#include "engine.h"
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cfm {

void* tensor_ptr(const Engine& e, int id, int B);

// ------------------------------------------------------------------------------------------------
// GroupNorm32 over NHWC bf16, register-resident.
// An "item" is one (sample, channel slab); a slab is a whole number of groups and of 8-channel
// 16-byte vectors (>= 64 B per pixel: 32 or 48 channels).  TPI threads own an item: each thread
// loads up to GN_VPT vectors straight into registers (all loads in flight at once - the kernel is
// HBM-bound, so bytes in flight are what matters), statistics are reduced deterministically
// through shared memory (no atomics -> bit-reproducible), and the normalised / FiLM-modulated /
// SiLU-activated values are stored from the same registers: one global read, one global write.
// Small feature maps pack several items into one CTA.
// ------------------------------------------------------------------------------------------------
struct GnFastArgs {
  const bf16* src0; const bf16* src1; int C0, C1;
  int HW, cpg, slab;                 // slab channels per item (multiple of 8 and of cpg)
  int tpi, ipc, n_items;             // threads per item, items per CTA, total items
  int cs;                            // CTAs (of one cluster) sharing an item's pixels; 1 = no cluster
  const float* gamma; const float* beta; float eps; int silu;
  const float* film; int film_stride; const int* film_row;
  bf16* out;
};

constexpr int GN_VPT = 16;
constexpr int GN_MAX_SLAB = 64;
constexpr int GN_MAX_THREADS = 256;

__global__ void __launch_bounds__(GN_MAX_THREADS, 2) groupnorm_bf16_kernel(GnFastArgs a) {
  extern __shared__ float gn_smem[];
  // layout: part_sum[threads][8] | part_sq[threads][8] | ch_scale[ipc][64] | ch_shift[ipc][64] | ch_sum[ipc][64] | ch_sq[ipc][64]
  float* part_sum = gn_smem;
  float* part_sq = part_sum + blockDim.x * 8;
  float* ch_scale = part_sq + blockDim.x * 8;
  float* ch_shift = ch_scale + a.ipc * GN_MAX_SLAB;
  float* ch_sum = ch_shift + a.ipc * GN_MAX_SLAB;
  float* ch_sq = ch_sum + a.ipc * GN_MAX_SLAB;

  const int C = a.C0 + a.C1;
  const int slabs = C / a.slab;
  const int vpp = a.slab / 8;
  const int nvec = a.HW * vpp;
  const int il = threadIdx.x / a.tpi;                    // item within the CTA
  const int ti = threadIdx.x - il * a.tpi;               // thread within the item
  // cs > 1: the CTAs of a cluster split one item's pixels (large feature maps); then ipc == 1
  const int crank = a.cs > 1 ? (int)cluster_ctarank() : 0;
  const int item = a.cs > 1 ? (int)(blockIdx.x / a.cs) : blockIdx.x * a.ipc + il;
  const int HWl = a.HW / a.cs;                           // pixels owned by this CTA
  const int pbase = crank * HWl;
  const bool active = item < a.n_items;
  const int b = active ? item / slabs : 0, sl = active ? item % slabs : 0;
  const int c_base = sl * a.slab;
  const int q = ti % vpp;                                // tpi % vpp == 0 -> fixed vector slot per thread
  const int cq = c_base + q * 8;
  const bf16* sp; int sC, sc;
  if (cq < a.C0) { sp = a.src0; sC = a.C0; sc = cq; } else { sp = a.src1; sC = a.C1; sc = cq - a.C0; }
  const long long pix0 = (long long)b * a.HW + pbase;
  const int p0 = ti / vpp, pstep = a.tpi / vpp;          // this thread's pixels: p0 + k*pstep (within the CTA's range)

  uint4 regs[GN_VPT];
  {
    const bf16* lp = sp + (pix0 + p0) * sC + sc;
    const long long lstep = (long long)pstep * sC;
#pragma unroll
    for (int k = 0; k < GN_VPT; ++k) {
      regs[k] = (active && p0 + k * pstep < HWl) ? __ldg((const uint4*)lp) : make_uint4(0, 0, 0, 0);
      lp += lstep;
    }
  }
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; ss[j] = 0.f; }
#pragma unroll
  for (int k = 0; k < GN_VPT; ++k) {
    const __nv_bfloat162* h2 = (const __nv_bfloat162*)&regs[k];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(h2[j]);
      s[2 * j] += f.x; ss[2 * j] = fmaf(f.x, f.x, ss[2 * j]);
      s[2 * j + 1] += f.y; ss[2 * j + 1] = fmaf(f.y, f.y, ss[2 * j + 1]);
    }
  }
  // keep only the packed bf16 words live across the reduction (stops the compiler from parking 128 unpacked floats)
#pragma unroll
  for (int k = 0; k < GN_VPT; ++k) asm volatile("" : "+r"(regs[k].x), "+r"(regs[k].y), "+r"(regs[k].z), "+r"(regs[k].w));
#pragma unroll
  for (int j = 0; j < 8; ++j) { part_sum[threadIdx.x * 8 + j] = s[j]; part_sq[threadIdx.x * 8 + j] = ss[j]; }
  __syncthreads();
  // one thread per (item, channel): fixed-order sum over the item's threads that share the vector slot
  for (int w = threadIdx.x; w < a.ipc * a.slab; w += blockDim.x) {
    const int wi = w / a.slab, c = w - wi * a.slab;
    const int qq = c >> 3, j = c & 7;
    float ts = 0.f, tq = 0.f;
    for (int t = wi * a.tpi + qq; t < (wi + 1) * a.tpi; t += vpp) { ts += part_sum[t * 8 + j]; tq += part_sq[t * 8 + j]; }
    ch_sum[wi * GN_MAX_SLAB + c] = ts; ch_sq[wi * GN_MAX_SLAB + c] = tq;
  }
  __syncthreads();
  if (a.cs > 1) {
    // combine the per-CTA channel sums across the cluster through distributed shared memory, in rank order
    cluster_sync_all();
    float ts = 0.f, tq = 0.f;
    if (threadIdx.x < a.slab) {
      for (int r = 0; r < a.cs; ++r) {
        ts += ld_dsmem_f32(mapa_u32(smem_u32(&ch_sum[threadIdx.x]), r));
        tq += ld_dsmem_f32(mapa_u32(smem_u32(&ch_sq[threadIdx.x]), r));
      }
    }
    cluster_sync_all();              // everyone has read the partials before they are overwritten
    if (threadIdx.x < a.slab) { ch_sum[threadIdx.x] = ts; ch_sq[threadIdx.x] = tq; }
    __syncthreads();
  }
  for (int w = threadIdx.x; w < a.ipc * a.slab; w += blockDim.x) {
    const int wi = w / a.slab, c = w - wi * a.slab;
    const int it = a.cs > 1 ? (int)(blockIdx.x / a.cs) : blockIdx.x * a.ipc + wi;
    if (it >= a.n_items) continue;
    const int bb = it / slabs, cb = (it % slabs) * a.slab;
    const int g0 = (c / a.cpg) * a.cpg;
    float gs = 0.f, gq = 0.f;
    for (int j = 0; j < a.cpg; ++j) { gs += ch_sum[wi * GN_MAX_SLAB + g0 + j]; gq += ch_sq[wi * GN_MAX_SLAB + g0 + j]; }
    const float inv_n = 1.0f / (float)(a.cpg * a.HW);
    const float mean = gs * inv_n;
    const float var = fmaxf(gq * inv_n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + a.eps);
    float sc_ = rstd * a.gamma[cb + c];
    float sh_ = a.beta[cb + c] - mean * sc_;
    if (a.film) {
      const float* f = a.film + (long long)a.film_row[bb] * a.film_stride;
      const float m = 1.0f + f[cb + c];
      sc_ *= m; sh_ = sh_ * m + f[C + cb + c];
    }
    if (a.silu) { sc_ *= 0.5f; sh_ *= 0.5f; }          // the activation works on h = y/2: silu(y) = h*tanh(h) + h
    ch_scale[wi * GN_MAX_SLAB + c] = sc_; ch_shift[wi * GN_MAX_SLAB + c] = sh_;
  }
  __syncthreads();
  if (!active) return;
  float sc8[8], sh8[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc8[j] = ch_scale[il * GN_MAX_SLAB + q * 8 + j]; sh8[j] = ch_shift[il * GN_MAX_SLAB + q * 8 + j]; }
  bf16* op = a.out + (pix0 + p0) * C + cq;
  const long long ostep = (long long)pstep * C;
#pragma unroll
  for (int k = 0; k < GN_VPT; ++k, op += ostep) {
    if (p0 + k * pstep < HWl) {
      const __nv_bfloat162* h2 = (const __nv_bfloat162*)&regs[k];
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
}

static int gcd_i(int a, int b) { return b ? gcd_i(b, a % b) : a; }

struct GnGeom { int cpg, slab, tpi, ipc, threads, cs; size_t smem; bool ok; };

static GnGeom gn_geometry(const Op& op) {
  GnGeom g{};
  const int C = op.Cin, HW = op.Hin * op.Win;
  g.cpg = C / 32;
  g.cs = 1;
  const int base = g.cpg / gcd_i(g.cpg, 8) * 8;        // lcm(cpg, 8): smallest legal slab
  // widest slab (<= 64 channels, dividing C) whose item still fits GN_VPT vectors/thread in <= 256 threads
  static const int max_slab = [] { const char* v = getenv("CFM_GN_MAXSLAB"); return v ? atoi(v) : GN_MAX_SLAB; }();
  static const int cta_threads = [] { const char* v = getenv("CFM_GN_CTA"); return v ? atoi(v) : GN_MAX_THREADS; }();
  for (int sl = base; sl <= std::max(max_slab, base) && sl <= GN_MAX_SLAB && C % sl == 0; sl *= 2) {
    const int vpp = sl / 8;
    const int unit = 32 / gcd_i(32, vpp) * vpp;        // lcm(32, vpp): whole warps, multiple of vpp
    const int nvec = HW * vpp;
    int tpi = ((nvec + GN_VPT - 1) / GN_VPT + unit - 1) / unit * unit;
    tpi = std::max(tpi, unit);
    if (tpi > GN_MAX_THREADS) break;
    g.slab = sl; g.tpi = tpi; g.ok = true;
  }
  static const bool wide = [] { const char* v = getenv("CFM_GN_WIDE"); return v && v[0] == '1'; }();
  int wbase = base;
  if (wide && g.ok && g.slab < 64 && C % 64 == 0 && 64 % g.cpg == 0 && HW >= 256) { g.ok = false; wbase = 64; }   // experiment: full 128 B lines per CTA
  if (!g.ok && wbase <= GN_MAX_SLAB && C % wbase == 0) {
    // large feature map: split the item's pixels over a cluster of 2..8 CTAs (partials combined through DSMEM)
    const int base = wbase;
    const int vpp = base / 8;
    const int unit = 32 / gcd_i(32, vpp) * vpp;
    for (int cs = 2; cs <= 8; cs *= 2) {
      if (HW % cs) break;
      const int nvec = (HW / cs) * vpp;
      int tpi = ((nvec + GN_VPT - 1) / GN_VPT + unit - 1) / unit * unit;
      tpi = std::max(tpi, unit);
      if (tpi <= GN_MAX_THREADS) { g.slab = base; g.tpi = tpi; g.cs = cs; g.ok = true; break; }
    }
  }
  if (!g.ok) return g;
  g.ipc = g.cs > 1 ? 1 : std::max(1, cta_threads / g.tpi);
  g.threads = g.tpi * g.ipc;
  g.smem = sizeof(float) * ((size_t)g.threads * 16 + (size_t)g.ipc * GN_MAX_SLAB * 4);
  return g;
}

bool gn_bf16_supported(const Engine& e, const Op& op) {
  if (!e.bf16 || op.kind != OP_GN) return false;
  const char* off = getenv("CFM_DISABLE_FAST_GN");
  if (off && off[0] == '1') return false;
  const GnGeom g = gn_geometry(op);
  if (!g.ok || g.smem > 96 * 1024) return false;
  if (e.tensors[op.src0].C % 8) return false;
  return true;
}

int gn_bf16_launch(Engine& e, const Op& op, int B, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(groupnorm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess) {
      e.err = "cudaFuncSetAttribute(groupnorm_bf16_kernel) failed"; return CFM_ERR_CUDA;
    }
    attr = true;
  }
  const GnGeom g = gn_geometry(op);
  GnFastArgs a{};
  a.cpg = g.cpg; a.slab = g.slab; a.tpi = g.tpi; a.ipc = g.ipc; a.cs = g.cs;
  a.src0 = (const bf16*)tensor_ptr(e, op.src0, B); a.C0 = e.tensors[op.src0].C;
  a.src1 = (const bf16*)tensor_ptr(e, op.src1, B); a.C1 = op.src1 >= 0 ? e.tensors[op.src1].C : 0;
  a.HW = op.Hin * op.Win;
  a.n_items = B * (op.Cin / g.slab);
  a.gamma = op.gamma; a.beta = op.beta; a.eps = 1e-5f; a.silu = op.silu;
  if (op.film) { a.film = e.emb_out + op.emb_off; a.film_stride = e.emb_total; a.film_row = e.row_of_sample; }
  a.out = (bf16*)tensor_ptr(e, op.out, B);
  if (g.cs > 1) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(a.n_items * g.cs));
    cfg.blockDim = dim3(g.threads);
    cfg.dynamicSmemBytes = g.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = g.cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, groupnorm_bf16_kernel, a) != cudaSuccess) { e.err = "groupnorm cluster launch failed"; return CFM_ERR_CUDA; }
    return 0;
  }
  const int blocks = (a.n_items + g.ipc - 1) / g.ipc;
  groupnorm_bf16_kernel<<<blocks, g.threads, g.smem, st>>>(a);
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
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(head_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess) { e.err = "cudaFuncSetAttribute(head_conv_kernel) failed"; return CFM_ERR_CUDA; }
    attr = true;
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
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess) { e.err = "cudaFuncSetAttribute(stem_conv_kernel) failed"; return CFM_ERR_CUDA; }
    attr = true;
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
__global__ void stem_im2col_kernel(const float* __restrict__ x0, const float* __restrict__ x1, int C0, int C1,
                                   bf16* __restrict__ out, int B, int H, int W, int Kpad) {
  const int Cin = C0 + C1, K9 = 9 * Cin;
  const int cv = Kpad >> 3;
  const long long total = (long long)B * H * W * cv;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int k0 = (int)(i % cv) << 3;
    const long long m = i / cv;
    const int x = (int)(m % W), y = (int)((m / W) % H), b = (int)(m / ((long long)W * H));
    uint4 o4;
    bf16* ob = (bf16*)&o4;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int k = k0 + j;
      const bool lo = k >= K9;
      if (lo) k -= K9;
      float v = 0.f;
      if (k < K9) {
        const int tap = k / Cin, c = k - tap * Cin;
        const int iy = y + tap / 3 - 1, ix = x + tap % 3 - 1;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W)
          v = (c < C0) ? __ldg(x0 + (((long long)b * C0 + c) * H + iy) * W + ix)
                       : __ldg(x1 + (((long long)b * C1 + (c - C0)) * H + iy) * W + ix);
        if (lo) v = v - __bfloat162float(__float2bfloat16_rn(v));
      }
      ob[j] = __float2bfloat16_rn(v);
    }
    *(uint4*)(out + i * 8) = o4;
  }
}

int stem_im2col_launch(Engine& e, const Op& op, int B, const float* x, const float* cond, cudaStream_t st) {
  const int cx = cond ? e.x_channels() : e.cfg.in_channels;
  const long long total = (long long)B * op.Hin * op.Win * (op.Cout / 8);
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)e.sm_count * 16);
  stem_im2col_kernel<<<blocks, 256, 0, st>>>(x, cond, cx, e.cfg.in_channels - cx, (bf16*)tensor_ptr(e, op.out, B), B, op.Hin, op.Win, op.Cout);
  return 0;
}

}  // namespace cfm
