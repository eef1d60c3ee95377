// Memory-bound bf16 kernels of the tensor-core path: GroupNorm(+FiLM)(+SiLU), nearest/avg resample,
// the 3-channel edge convolutions.  All are HBM-bound: 16-byte vector accesses, each tensor read once.
#include <algorithm>
#include "engine.h"
#include "common.cuh"

namespace cfm {

void* tensor_ptr(const Engine& e, int id, int B);

// ------------------------------------------------------------------------------------------------
// GroupNorm32 over NHWC bf16.  One CTA per (sample, channel slab); the slab (a whole number of groups
// and of 8-channel vectors, >= 64 B per pixel) for all HW pixels is staged once in shared memory,
// statistics are reduced from there (warp shuffles + a few shared atomics) and the normalised,
// FiLM-modulated, SiLU-activated result is written back with 16-byte stores.
// ------------------------------------------------------------------------------------------------
struct GnFastArgs {
  const bf16* src0; const bf16* src1; int C0, C1;
  int HW, cpg, slab;                 // slab channels per CTA (multiple of 8 and of cpg)
  const float* gamma; const float* beta; float eps; int silu;
  const float* film; int film_stride; const int* film_row;
  bf16* out;
};

constexpr int GN_THREADS = 256;
constexpr int GN_MAX_SLAB = 64;

__global__ void __launch_bounds__(GN_THREADS) groupnorm_bf16_kernel(GnFastArgs a) {
  extern __shared__ uint4 stage[];                       // [HW][slab/8] vectors
  __shared__ float ch_sum[GN_MAX_SLAB], ch_sq[GN_MAX_SLAB];
  __shared__ float ch_scale[GN_MAX_SLAB], ch_shift[GN_MAX_SLAB];
  __shared__ float part_sum[GN_THREADS][8], part_sq[GN_THREADS][8];
  const int C = a.C0 + a.C1;
  const int slabs = C / a.slab;
  const int b = blockIdx.x / slabs, sl = blockIdx.x % slabs;
  const int c_base = sl * a.slab;
  const int vpp = a.slab / 8;                            // 16-byte vectors per pixel
  const int nvec = a.HW * vpp;

  // GN_THREADS % vpp == 0 is guaranteed by the launcher (vpp in {3,4,6,8}: blockDim is chosen), so a
  // thread always sees the same vector slot q -> the same 8 channels.
  const int q = threadIdx.x % vpp;
  const int cq = c_base + q * 8;                         // first channel of this thread's vector
  const bf16* sp; int sC, sc;
  if (cq < a.C0) { sp = a.src0; sC = a.C0; sc = cq; } else { sp = a.src1; sC = a.C1; sc = cq - a.C0; }
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; ss[j] = 0.f; }
  const long long pix0 = (long long)b * a.HW;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    const int p = v / vpp;
    const uint4 raw = __ldg((const uint4*)(sp + (pix0 + p) * sC + sc));
    stage[v] = raw;
    const __nv_bfloat162* h2 = (const __nv_bfloat162*)&raw;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(h2[j]);
      s[2 * j] += f.x; ss[2 * j] = fmaf(f.x, f.x, ss[2 * j]);
      s[2 * j + 1] += f.y; ss[2 * j + 1] = fmaf(f.y, f.y, ss[2 * j + 1]);
    }
  }
  // deterministic reduction: every thread parks its 8 channel partials, then one thread per channel
  // adds the partials of the threads sharing its vector slot q in a fixed order (no atomics).
#pragma unroll
  for (int j = 0; j < 8; ++j) { part_sum[threadIdx.x][j] = s[j]; part_sq[threadIdx.x][j] = ss[j]; }
  __syncthreads();
  if (threadIdx.x < a.slab) {
    const int qq = threadIdx.x >> 3, j = threadIdx.x & 7;
    float ts = 0.f, tq = 0.f;
    for (int t = qq; t < (int)blockDim.x; t += vpp) { ts += part_sum[t][j]; tq += part_sq[t][j]; }
    ch_sum[threadIdx.x] = ts; ch_sq[threadIdx.x] = tq;
  }
  __syncthreads();
  if (threadIdx.x < a.slab) {
    const int c = threadIdx.x, g0 = (c / a.cpg) * a.cpg;
    float gs = 0.f, gq = 0.f;
    for (int j = 0; j < a.cpg; ++j) { gs += ch_sum[g0 + j]; gq += ch_sq[g0 + j]; }
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
    ch_scale[c] = sc_; ch_shift[c] = sh_;
  }
  __syncthreads();
  float sc8[8], sh8[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc8[j] = ch_scale[q * 8 + j]; sh8[j] = ch_shift[q * 8 + j]; }
  bf16* op = a.out + pix0 * C + cq;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    const int p = v / vpp;
    const uint4 raw = stage[v];
    const __nv_bfloat162* h2 = (const __nv_bfloat162*)&raw;
    uint4 o4;
    __nv_bfloat162* o2 = (__nv_bfloat162*)&o4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(h2[j]);
      float y0 = fmaf(f.x, sc8[2 * j], sh8[2 * j]), y1 = fmaf(f.y, sc8[2 * j + 1], sh8[2 * j + 1]);
      if (a.silu) { y0 = silu_f(y0); y1 = silu_f(y1); }
      o2[j] = __floats2bfloat162_rn(y0, y1);
    }
    *(uint4*)(op + (long long)p * C) = o4;
  }
}

static int gcd_i(int a, int b) { return b ? gcd_i(b, a % b) : a; }

static void gn_geometry(const Engine& e, const Op& op, int* cpg, int* slab, int* threads, size_t* smem) {
  const int C = op.Cin;
  *cpg = C / 32;
  int base = *cpg / gcd_i(*cpg, 8) * 8;            // lcm(cpg, 8)
  int sl = base;
  while (sl < 32 && C % (sl * 2) == 0) sl *= 2;
  *slab = sl;
  const int vpp = sl / 8;
  *threads = (GN_THREADS / vpp) * vpp;
  *smem = (size_t)op.Hin * op.Win * sl * 2;
}

bool gn_bf16_supported(const Engine& e, const Op& op) {
  if (!e.bf16 || op.kind != OP_GN) return false;
  const char* off = getenv("CFM_DISABLE_FAST_GN");
  if (off && off[0] == '1') return false;
  int cpg, slab, threads; size_t smem;
  gn_geometry(e, op, &cpg, &slab, &threads, &smem);
  if (slab > GN_MAX_SLAB || op.Cin % slab) return false;
  if (smem > 200 * 1024) return false;
  const int C0 = e.tensors[op.src0].C;
  if (C0 % 8) return false;
  return true;
}

int gn_bf16_launch(Engine& e, const Op& op, int B, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(groupnorm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
      e.err = "cudaFuncSetAttribute(groupnorm_bf16_kernel) failed"; return CFM_ERR_CUDA;
    }
    attr = true;
  }
  GnFastArgs a{};
  int threads; size_t smem;
  gn_geometry(e, op, &a.cpg, &a.slab, &threads, &smem);
  a.src0 = (const bf16*)tensor_ptr(e, op.src0, B); a.C0 = e.tensors[op.src0].C;
  a.src1 = (const bf16*)tensor_ptr(e, op.src1, B); a.C1 = op.src1 >= 0 ? e.tensors[op.src1].C : 0;
  a.HW = op.Hin * op.Win;
  a.gamma = op.gamma; a.beta = op.beta; a.eps = 1e-5f; a.silu = op.silu;
  if (op.film) { a.film = e.emb_out + op.emb_off; a.film_stride = e.emb_total; a.film_row = e.row_of_sample; }
  a.out = (bf16*)tensor_ptr(e, op.out, B);
  const int blocks = B * (op.Cin / a.slab);
  groupnorm_bf16_kernel<<<blocks, threads, smem, st>>>(a);
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

__global__ void __launch_bounds__(256) stem_conv_kernel(const float* __restrict__ x0, const float* __restrict__ x1, int C0, int C1,
                                                        const float* __restrict__ w /*[9*Cin][Cout]*/, const float* __restrict__ bias,
                                                        bf16* __restrict__ out, int B, int H, int W, int Cout) {
  const int Cin = C0 + C1;
  const int cv = Cout / 8;
  const long long total = (long long)B * H * W * cv;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int co = (int)(i % cv) * 8;
  const long long m = i / cv;
  const int x = (int)(m % W), y = (int)((m / W) % H), b = (int)(m / ((long long)W * H));
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = bias[co + j];
  for (int tap = 0; tap < 9; ++tap) {
    const int iy = y + tap / 3 - 1, ix = x + tap % 3 - 1;
    if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
    for (int c = 0; c < Cin; ++c) {
      const float v = (c < C0) ? __ldg(x0 + (((long long)b * C0 + c) * H + iy) * W + ix)
                               : __ldg(x1 + (((long long)b * C1 + (c - C0)) * H + iy) * W + ix);
      const float4 w0 = __ldg((const float4*)(w + (long long)(tap * Cin + c) * Cout + co));
      const float4 w1 = __ldg((const float4*)(w + (long long)(tap * Cin + c) * Cout + co + 4));
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

bool stem_conv_supported(const Engine& e, const Op& op) {
  return e.bf16 && op.kind == OP_CONV && op.src_is_input && !op.out_is_output && op.ks == 3 && op.stride == 1 && !op.ups &&
         op.skip0 < 0 && op.res0 < 0 && op.emb_off < 0 && op.Cin <= STEM_MAX_CIN && op.Cout % 8 == 0;
}

int stem_conv_launch(Engine& e, const Op& op, int B, const float* x, const float* cond, cudaStream_t st) {
  const int cx = cond ? e.x_channels() : e.cfg.in_channels;
  const long long total = (long long)B * op.Hout * op.Wout * (op.Cout / 8);
  stem_conv_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, cond, cx, e.cfg.in_channels - cx, op.w_main, op.bias,
                                                                  (bf16*)tensor_ptr(e, op.out, B), B, op.Hout, op.Wout, op.Cout);
  return 0;
}

}  // namespace cfm
