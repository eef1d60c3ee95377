// Fused integrator / reverse-chain step kernels (memory-bound, one launch per step).
//
// Arithmetic uses explicitly rounded multiplies/adds (no FMA contraction) in the order the
// reference's PyTorch expressions evaluate them, so that, given the same U-Net output and the
// same injected noise, a step is bit-identical to the reference's CPU elementwise result.
#pragma once
#include "common.cuh"

namespace cfm {

// --- counter-based normal generator (Philox4x32-10 + Box-Muller) ---------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * ctr.x;
    const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * ctr.z;
    uint4 n;
    n.x = (unsigned)(p1 >> 32) ^ ctr.y ^ key.x;
    n.y = (unsigned)p1;
    n.z = (unsigned)(p0 >> 32) ^ ctr.w ^ key.y;
    n.w = (unsigned)p0;
    ctr = n;
    key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
  }
  return ctr;
}
// one standard normal for (seed, stream, element index)
__device__ __forceinline__ float philox_normal(unsigned long long seed, unsigned stream, unsigned long long idx) {
  const uint4 r = philox4x32_10(make_uint4((unsigned)(idx >> 1), (unsigned)(idx >> 33), stream, 0u),
                                make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
  const float u1 = ((r.x >> 8) + 1u) * (1.0f / 16777216.0f);     // (0, 1]
  const float u2 = (r.y >> 8) * (1.0f / 16777216.0f);            // [0, 1)
  const float rad = sqrtf(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.28318530717958647692f * u2, &s, &c);
  return (idx & 1ull) ? rad * s : rad * c;
}

// Per-call values that must not be baked into a captured step graph: the injected-noise tensor, the Philox seed and the
// caller's trajectory buffer.  They live in device memory (Engine::sampler_params) and are rewritten before every loop, so
// one cached graph serves any noise tensor, seed and trajectory buffer.
struct SamplerParams { const float* noise; unsigned long long seed; float* traj; };

__device__ __forceinline__ uint8_t quant_u8(float x) { return (uint8_t)fminf(fmaxf(__fadd_rn(__fmul_rn(x, 127.5f), 128.0f), 0.0f), 255.0f); }

// --- Euler:  x <- x + dt * v   (torchdyn fixed-step driver) ----------------------------------------
// cond (optional, COND_DRIFT): con <- con + dt * con  (the reference's ode_func returns x[1] as d(con)/dt)
// traj (optional, from SamplerParams): next trajectory slot receives the new x.  img (optional): uint8 quantisation.
// Main loop in 16-byte vectors (n4 = n / 4 of them; x, v, traj slots and img are 16-byte aligned: cudaMalloc bases and
// n % 4 == 0 offsets), scalar tail.  HBM-bound: 2 reads + 1 write (+ 1 per optional output) of 4 B per element.
__global__ void euler_step_kernel(float* __restrict__ x, const float* __restrict__ v, const float* __restrict__ dt_table,
                                  const int* __restrict__ step_counter, int n_steps, long long n,
                                  float* __restrict__ cond, long long n_cond,
                                  const SamplerParams* __restrict__ sp, uint8_t* __restrict__ img) {
  const int k = *step_counter;
  const float dt = dt_table[k];
  float* traj = sp ? sp->traj : nullptr;
  float* traj_slot = traj ? traj + (long long)(k + 1) * n : nullptr;
  const bool want_img = img && k == n_steps - 1;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = (n & 3) == 0 && (((uintptr_t)traj & 15) == 0);
  const long long n4 = vec ? n >> 2 : 0;
  for (long long i = t0; i < n4; i += stride) {
    const float4 xv = ((const float4*)x)[i], vv = ((const float4*)v)[i];
    float4 o;
    o.x = __fadd_rn(xv.x, __fmul_rn(dt, vv.x)); o.y = __fadd_rn(xv.y, __fmul_rn(dt, vv.y));
    o.z = __fadd_rn(xv.z, __fmul_rn(dt, vv.z)); o.w = __fadd_rn(xv.w, __fmul_rn(dt, vv.w));
    ((float4*)x)[i] = o;
    if (traj_slot) ((float4*)traj_slot)[i] = o;
    if (want_img) ((uchar4*)img)[i] = make_uchar4(quant_u8(o.x), quant_u8(o.y), quant_u8(o.z), quant_u8(o.w));
  }
  for (long long i = n4 * 4 + t0; i < n; i += stride) {
    const float nx = __fadd_rn(x[i], __fmul_rn(dt, v[i]));
    x[i] = nx;
    if (traj_slot) traj_slot[i] = nx;
    if (want_img) img[i] = quant_u8(nx);
  }
  if (cond)
    for (long long i = t0; i < n_cond; i += stride)
      cond[i] = __fadd_rn(cond[i], __fmul_rn(dt, cond[i]));
}

__global__ void fill_f32_kernel(float* __restrict__ p, float v, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

// classifier-free guidance: v_c <- v_c + w (v_c - v_u)   (= (1 + w) v_c - w v_u), evaluated in this order
__global__ void cfg_combine_kernel(float* __restrict__ vc, const float* __restrict__ vu, float w, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    vc[i] = __fadd_rn(vc[i], __fmul_rn(w, __fsub_rn(vc[i], vu[i])));
}

__global__ void quantize_u8_kernel(uint8_t* __restrict__ out, const float* __restrict__ x, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = (n & 3) == 0 && (((uintptr_t)out & 3) == 0) && (((uintptr_t)x & 15) == 0);
  const long long n4 = vec ? n >> 2 : 0;
  for (long long i = t0; i < n4; i += stride) {
    const float4 v = ((const float4*)x)[i];
    ((uchar4*)out)[i] = make_uchar4(quant_u8(v.x), quant_u8(v.y), quant_u8(v.z), quant_u8(v.w));
  }
  for (long long i = n4 * 4 + t0; i < n; i += stride) out[i] = quant_u8(x[i]);
}

// --- DDPM ---------------------------------------------------------------------------------------
struct DdpmStepScalars {
  // posterior at step i
  float a, b, c1, c2, sigma;   // sqrt_recip, sqrt_recipm1, coef1, coef2, exp(0.5*logvar)
  int add_noise;               // i > 0
  // blend for the NEXT step (i-1), applied to the freshly computed x
  int blend_next; int noise_condition; float sa, sb, pad_value;   // sqrt_ac[i-1], sqrt_1mac[i-1]
  int final_clip;              // i == 0: clip(x, -1, 1)
  int chain_index;             // i (noise slots / Philox streams are indexed by it)
  // Langevin corrector steps (sampling.py:241-250): the blend of step i then runs as its own launch BEFORE the U-Net
  // call (blend_cur) because corrector evaluations sit between the posterior draw and the next step's blend
  int n_slots;                 // noise slots per chain step: 2 + n_corrector
  int blend_cur; float sa_cur, sb_cur;   // sqrt_ac[i], sqrt_1mac[i]
  float corr_r, corr_cd, corr_cn;        // 1/sqrt(1-abar_i), 0.5*dt*delta, sqrt(dt*delta)
};

// x: in = xi (already blended for step i), out = state handed to the next U-Net call.
// noise (optional): [Ns, n_slots, n]; slot (i,0) = blend draw, (i,1) = posterior draw; without it a Philox generator is
// used (streams n_slots*i / n_slots*i + 1).
__device__ __forceinline__ void ddpm_step_body(float* __restrict__ x, const float* __restrict__ eps, const DdpmStepScalars& s,
                                               const float* __restrict__ cond, const float* __restrict__ noise,
                                               unsigned long long seed, long long n) {
  const int ci = s.chain_index;
  const int ns = s.n_slots;
  const float* z_post = noise ? noise + ((long long)ci * ns + 1) * n : nullptr;
  const float* z_blend = (noise && s.blend_next) ? noise + ((long long)(ci - 1) * ns) * n : nullptr;
  const unsigned stream_post = (unsigned)(ns * ci + 1), stream_blend = (unsigned)(ns * (ci - 1));
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float xi = x[i];
    float x0 = __fsub_rn(__fmul_rn(s.a, xi), __fmul_rn(s.b, eps[i]));
    x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
    float nx = __fadd_rn(__fmul_rn(s.c1, x0), __fmul_rn(s.c2, xi));
    if (s.add_noise) {
      const float z = z_post ? z_post[i] : philox_normal(seed, stream_post, (unsigned long long)i);
      nx = __fadd_rn(nx, __fmul_rn(s.sigma, z));
    }
    if (s.blend_next) {
      const float c = cond[i];
      float nc = c;
      if (s.noise_condition) {
        const float z = z_blend ? z_blend[i] : philox_normal(seed, stream_blend, (unsigned long long)i);
        nc = __fadd_rn(__fmul_rn(s.sa, c), __fmul_rn(s.sb, z));
      }
      nx = (c == s.pad_value) ? nx : nc;
    }
    if (s.final_clip) nx = fminf(fmaxf(nx, -1.0f), 1.0f);
    x[i] = nx;
  }
}
// Blend of the CURRENT step (chains with corrector steps, stepwise API): x = where(cond == pad, x, q_sample(cond)).
__device__ __forceinline__ void ddpm_blend_cur_body(float* __restrict__ x, const float* __restrict__ cond, const DdpmStepScalars& s,
                                                    const float* __restrict__ noise, unsigned long long seed, long long n) {
  if (!s.blend_cur) return;
  const float* z_blend = noise ? noise + ((long long)s.chain_index * s.n_slots) * n : nullptr;
  const unsigned stream = (unsigned)(s.n_slots * s.chain_index);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float c = cond[i];
    float nc = c;
    if (s.noise_condition) {
      const float z = z_blend ? z_blend[i] : philox_normal(seed, stream, (unsigned long long)i);
      nc = __fadd_rn(__fmul_rn(s.sa_cur, c), __fmul_rn(s.sb_cur, z));
    }
    if (!(c == s.pad_value)) x[i] = nc;
  }
}
// Langevin corrector (sampling.py:241-250, sde_diffusion.py:214-217): with eps = model(x, t_i),
//   x0 = clip(a x - b eps),  score = -(x0 / sqrt(1 - abar_i)),  x += (0.5 dt delta) score + sqrt(dt delta) z
// in the reference's order of fp32 operations.  `last`: final corrector of chain step 0 -> clip(x, -1, 1).
__device__ __forceinline__ void ddpm_corrector_body(float* __restrict__ x, const float* __restrict__ eps, const DdpmStepScalars& s,
                                                    const float* __restrict__ noise, unsigned long long seed, int slot, int last,
                                                    long long n) {
  const float* zc = noise ? noise + ((long long)s.chain_index * s.n_slots + slot) * n : nullptr;
  const unsigned stream = (unsigned)(s.n_slots * s.chain_index + slot);
  const bool clip = last && s.chain_index == 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float xi = x[i];
    float x0 = __fsub_rn(__fmul_rn(s.a, xi), __fmul_rn(s.b, eps[i]));
    x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
    const float score = -__fmul_rn(s.corr_r, x0);
    const float z = zc ? zc[i] : philox_normal(seed, stream, (unsigned long long)i);
    float nx = __fadd_rn(xi, __fadd_rn(__fmul_rn(s.corr_cd, score), __fmul_rn(s.corr_cn, z)));
    if (clip) nx = fminf(fmaxf(nx, -1.0f), 1.0f);
    x[i] = nx;
  }
}
// Fused-loop launches: the step's scalars come from a device table indexed by the device-side step counter and the
// noise pointer / seed from SamplerParams, so one captured launch serves every step of every call.
__global__ void ddpm_step_kernel(float* __restrict__ x, const float* __restrict__ eps,
                                 const DdpmStepScalars* __restrict__ table, const int* __restrict__ step_counter,
                                 const float* __restrict__ cond, const SamplerParams* __restrict__ sp, long long n) {
  const DdpmStepScalars s = table[*step_counter];
  ddpm_step_body(x, eps, s, cond, sp->noise, sp->seed, n);
}
__global__ void ddpm_blend_table_kernel(float* __restrict__ x, const float* __restrict__ cond,
                                        const DdpmStepScalars* __restrict__ table, const int* __restrict__ step_counter,
                                        const SamplerParams* __restrict__ sp, long long n) {
  const DdpmStepScalars s = table[*step_counter];
  ddpm_blend_cur_body(x, cond, s, sp->noise, sp->seed, n);
}
__global__ void ddpm_corrector_kernel(float* __restrict__ x, const float* __restrict__ eps,
                                      const DdpmStepScalars* __restrict__ table, const int* __restrict__ step_counter,
                                      const SamplerParams* __restrict__ sp, int slot, int last, long long n) {
  const DdpmStepScalars s = table[*step_counter];
  ddpm_corrector_body(x, eps, s, sp->noise, sp->seed, slot, last, n);
}
// Stepwise launches (cfm_ddpm_step: a Python-level eps network between the steps): scalars by value.
__global__ void ddpm_step_value_kernel(float* __restrict__ x, const float* __restrict__ eps, DdpmStepScalars s,
                                       const float* __restrict__ cond, const float* __restrict__ noise, unsigned long long seed, long long n) {
  ddpm_step_body(x, eps, s, cond, noise, seed, n);
}
__global__ void ddpm_blend_value_kernel(float* __restrict__ x, const float* __restrict__ cond, DdpmStepScalars s,
                                        const float* __restrict__ noise, unsigned long long seed, long long n) {
  ddpm_blend_cur_body(x, cond, s, noise, seed, n);
}
__global__ void ddpm_corrector_value_kernel(float* __restrict__ x, const float* __restrict__ eps, DdpmStepScalars s,
                                            const float* __restrict__ noise, unsigned long long seed, int slot, int last, long long n) {
  ddpm_corrector_body(x, eps, s, noise, seed, slot, last, n);
}

// Blend alone (before the first U-Net call of the chain): x = where(cond == pad, x, q_sample(cond)).
__global__ void ddpm_blend_kernel(float* __restrict__ x, const float* __restrict__ cond, float sa, float sb,
                                  float pad_value, int noise_condition, const float* __restrict__ z_blend,
                                  unsigned long long seed, unsigned stream_blend, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float c = cond[i];
    float nc = c;
    if (noise_condition) {
      const float z = z_blend ? z_blend[i] : philox_normal(seed, stream_blend, (unsigned long long)i);
      nc = __fadd_rn(__fmul_rn(sa, c), __fmul_rn(sb, z));
    }
    if (!(c == pad_value)) x[i] = nc;
  }
}

// --- Euler-Maruyama SDE steps --------------------------------------------------------------------
// SF2M sampling (conditional_mnist.ipynb cells 11-12: torchsde.sdeint of f = flow + score, g = sigma, fixed dt; the
// "euler" scheme of an Ito SDE with diagonal noise):  x <- x + (v + s) * dt + sigma * dW,  dW = sqrt(dt) * z.
// score may be NULL (plain drift).  z from `noise` (injected, n values) or Philox (seed, stream).
__global__ void sde_em_step_kernel(float* __restrict__ x, const float* __restrict__ drift, const float* __restrict__ score,
                                   float dt, float sigma, float sqrt_dt, const float* __restrict__ noise,
                                   unsigned long long seed, unsigned stream, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float f = score ? __fadd_rn(drift[i], score[i]) : drift[i];
    const float z = noise ? noise[i] : philox_normal(seed, stream, (unsigned long long)i);
    const float dw = __fmul_rn(z, sqrt_dt);
    x[i] = __fadd_rn(__fadd_rn(x[i], __fmul_rn(f, dt)), __fmul_rn(sigma, dw));
  }
}
// Reverse-time VP-SDE step of the Amortized sampler (sampling.py:100-111 `em_step`, sde_diffusion.py:170-205), in the
// reference's order of fp32 operations:  score = -eps / sigma_t,  drift = ((-0.5 x) x) - g^2 score  (g = sqrt(beta_t)),
//   x <- (x - dt * drift) + (g * z) * sqrt(dt)
// (the reference's DDPM.drift evaluates -0.5 * unsqueeze_like(beta_t, x) * x with the helper's arguments swapped, which
//  yields -0.5 * x * x; reproduced as is - parity is with what the reference computes)
__global__ void ddpm_em_step_kernel(float* __restrict__ x, const float* __restrict__ eps, float beta_t, float sigma_t, float g,
                                    float dt, float sqrt_dt, const float* __restrict__ noise, unsigned long long seed,
                                    unsigned stream, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float g2 = __fmul_rn(g, g);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float xi = x[i];
    const float score = -(eps[i] / sigma_t);
    const float drift = __fsub_rn(__fmul_rn(__fmul_rn(-0.5f, xi), xi), __fmul_rn(g2, score));
    const float z = noise ? noise[i] : philox_normal(seed, stream, (unsigned long long)i);
    x[i] = __fadd_rn(__fsub_rn(xi, __fmul_rn(dt, drift)), __fmul_rn(__fmul_rn(g, z), sqrt_dt));
  }
}

// --- bilinear resize (F.interpolate(mode="bilinear", align_corners=False), no antialias) -----------
// likelihoods.py:119-126 (HyperResolution: down then up), mnist/utils_mnist_hy.py:18-28 (downsample_images), and the
// low_res -> full-size upsample of SuperResModelWrapper.  ATen's source index: src = max(scale * (dst + 0.5) - 0.5, 0),
// scale = in / out; weights (1 - l, l); value = w0y * (w0x v00 + w1x v01) + w1y * (w0x v10 + w1x v11).  NCHW fp32.
__global__ void resize_bilinear_kernel(float* __restrict__ out, const float* __restrict__ in, long long planes, int Hin, int Win,
                                       int Hout, int Wout, float scale_h, float scale_w) {
  const long long n = planes * Hout * Wout;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int ox = (int)(i % Wout), oy = (int)((i / Wout) % Hout);
    const long long pl = i / ((long long)Wout * Hout);
    const float sy = fmaxf(__fsub_rn(__fmul_rn(scale_h, __fadd_rn((float)oy, 0.5f)), 0.5f), 0.f);
    const float sx = fmaxf(__fsub_rn(__fmul_rn(scale_w, __fadd_rn((float)ox, 0.5f)), 0.5f), 0.f);
    const int y0 = min((int)sy, Hin - 1), x0 = min((int)sx, Win - 1);
    const int y1 = min(y0 + 1, Hin - 1), x1 = min(x0 + 1, Win - 1);
    const float ly = fminf(fmaxf(__fsub_rn(sy, (float)y0), 0.f), 1.f), lx = fminf(fmaxf(__fsub_rn(sx, (float)x0), 0.f), 1.f);
    const float wy0 = __fsub_rn(1.f, ly), wx0 = __fsub_rn(1.f, lx);
    const float* p = in + pl * Hin * Win;
    const float top = __fadd_rn(__fmul_rn(wx0, p[y0 * Win + x0]), __fmul_rn(lx, p[y0 * Win + x1]));
    const float bot = __fadd_rn(__fmul_rn(wx0, p[y1 * Win + x0]), __fmul_rn(lx, p[y1 * Win + x1]));
    out[i] = __fadd_rn(__fmul_rn(wy0, top), __fmul_rn(ly, bot));
  }
}

// --- dopri5 state algebra ----------------------------------------------------------------------
struct RkPtrs { const float* k[8]; float coef[8]; int n_k; };

__global__ void rk_combine_kernel(float* __restrict__ out, const float* __restrict__ y, RkPtrs p, float dt, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < p.n_k) s = (j == 0) ? __fmul_rn(p.coef[0], p.k[0][i]) : __fadd_rn(s, __fmul_rn(p.coef[j], p.k[j][i]));
    out[i] = __fadd_rn(y[i], __fmul_rn(dt, s));
  }
}

// Deterministic grid total of one fp64 value per thread (shared by the dopri5 norms below): every block leaves its
// partial sum (fixed intra-block order) in partial[blockIdx.x]; the block that finishes LAST (a ticket counter decides who
// that is - the counter orders nothing arithmetic) adds the partials in block order and writes the result.
__device__ __forceinline__ void rk_block_total(double acc, double* __restrict__ sumsq, double* __restrict__ partial,
                                               unsigned* __restrict__ ticket) {
  __shared__ double red[32];
  __shared__ bool is_last;
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = acc;
  __syncthreads();
  if (w == 0) {
    acc = (lane < (int)(blockDim.x >> 5)) ? red[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      partial[blockIdx.x] = acc;
      __threadfence();
      is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    double tot = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) tot += ((volatile double*)partial)[b];
    *sumsq = tot;
    *ticket = 0u;
  }
}

// Squared error norm of dopri5, deterministic: every block leaves its partial sum (fp64, fixed intra-block order) in
// partial[blockIdx.x]; the block that finishes LAST (a ticket counter decides who that is - the counter orders nothing
// arithmetic) adds the partials in block order and writes the result.  The same inputs give the same bits on every run,
// so an accept / reject decision at ratio ~ 1 cannot flip between runs.
__global__ void rk_error_sumsq_kernel(double* __restrict__ sumsq, double* __restrict__ partial, unsigned* __restrict__ ticket,
                                      const float* __restrict__ y0, const float* __restrict__ y1,
                                      RkPtrs p, float dt, float rtol, float atol, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < p.n_k) s = (j == 0) ? __fmul_rn(p.coef[0], p.k[0][i]) : __fadd_rn(s, __fmul_rn(p.coef[j], p.k[j][i]));
    const float err = __fmul_rn(dt, s);
    const float tol = __fadd_rn(atol, __fmul_rn(rtol, fmaxf(fabsf(y0[i]), fabsf(y1[i]))));
    const float r = err / tol;
    acc += (double)r * (double)r;
  }
  rk_block_total(acc, sumsq, partial, ticket);
}

// Initial-step heuristic of dopri5 (torchdiffeq `_select_initial_step`): sum_i ((a_i - b_i) / (atol + rtol * |y_i|))^2
// (b may be NULL), reduced like the error norm above.
__global__ void rk_scaled_sumsq_kernel(double* __restrict__ sumsq, double* __restrict__ partial, unsigned* __restrict__ ticket,
                                       const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ y,
                                       float rtol, float atol, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float num = b ? __fsub_rn(a[i], b[i]) : a[i];
    const float r = num / __fadd_rn(atol, __fmul_rn(fabsf(y[i]), rtol));
    acc += (double)r * (double)r;
  }
  rk_block_total(acc, sumsq, partial, ticket);
}

// Dense output of an accepted dopri5 step at fraction x of the step (torchdiffeq `_interp_fit` + `_interp_evaluate`):
// the quartic through y0, y_mid, y1 with end slopes f0, f1, evaluated by Horner in the reference's operation order.
__global__ void rk_dense_output_kernel(float* __restrict__ out, const float* __restrict__ y0, const float* __restrict__ y1,
                                       const float* __restrict__ ym, const float* __restrict__ f0, const float* __restrict__ f1,
                                       float dt, float x, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float a0 = y0[i], a1 = y1[i], am = ym[i], fa = f0[i], fb = f1[i];
    // a = 2 dt (fb - fa) - 8 (y1 + y0) + 16 ym
    const float ca = __fadd_rn(__fsub_rn(__fmul_rn(__fmul_rn(2.f, dt), __fsub_rn(fb, fa)), __fmul_rn(8.f, __fadd_rn(a1, a0))), __fmul_rn(16.f, am));
    // b = dt (5 fa - 3 fb) + 18 y0 + 14 y1 - 32 ym
    const float cb = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(dt, __fsub_rn(__fmul_rn(5.f, fa), __fmul_rn(3.f, fb))), __fmul_rn(18.f, a0)), __fmul_rn(14.f, a1)), __fmul_rn(32.f, am));
    // c = dt (fb - 4 fa) - 11 y0 - 5 y1 + 16 ym
    const float cc = __fadd_rn(__fsub_rn(__fsub_rn(__fmul_rn(dt, __fsub_rn(fb, __fmul_rn(4.f, fa))), __fmul_rn(11.f, a0)), __fmul_rn(5.f, a1)), __fmul_rn(16.f, am));
    const float cd = __fmul_rn(dt, fa);
    float r = __fadd_rn(__fmul_rn(ca, x), cb);
    r = __fadd_rn(__fmul_rn(r, x), cc);
    r = __fadd_rn(__fmul_rn(r, x), cd);
    out[i] = __fadd_rn(__fmul_rn(r, x), a0);
  }
}

// --- FID sufficient statistics (cifar10/compute_fid.py:92-100, AD/experiments/main.py:261-267, 292-293) -----------
// sum[d] += sum_i f[i][d];  outer[d][e] += sum_i f[i][d] f[i][e]  in fp64, over n feature rows of width D (fp32 in).
// One 32 x 32 tile of `outer` per CTA (16 x 16 threads, 2 x 2 outputs each); every output element is accumulated by ONE
// thread in row order, so repeated runs give the same bits; the running totals make the kernel callable batch by batch.
__global__ void __launch_bounds__(256) fid_accumulate_kernel(double* __restrict__ sum, double* __restrict__ outer,
                                                             const float* __restrict__ f, long long n, int D) {
  __shared__ float sa[32][33], sb[32][33];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int d0 = blockIdx.y * 32, e0 = blockIdx.x * 32;
  double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  double colsum = 0.0;                      // blockIdx.y == 0 CTAs also own sum[e0 .. e0 + 32)
  for (long long r0 = 0; r0 < n; r0 += 32) {
    for (int i = threadIdx.x; i < 32 * 32; i += 256) {
      const int rr = i >> 5, cc = i & 31;
      const long long r = r0 + rr;
      sa[rr][cc] = (r < n && d0 + cc < D) ? f[r * D + d0 + cc] : 0.f;
      sb[rr][cc] = (r < n && e0 + cc < D) ? f[r * D + e0 + cc] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < 32; ++rr) {
      const double a0 = sa[rr][ty], a1 = sa[rr][ty + 16], b0 = sb[rr][tx], b1 = sb[rr][tx + 16];
      acc[0][0] = fma(a0, b0, acc[0][0]); acc[0][1] = fma(a0, b1, acc[0][1]);
      acc[1][0] = fma(a1, b0, acc[1][0]); acc[1][1] = fma(a1, b1, acc[1][1]);
    }
    if (blockIdx.y == 0 && threadIdx.x < 32)
      for (int rr = 0; rr < 32; ++rr) colsum += (double)sb[rr][threadIdx.x];
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int d = d0 + ty + 16 * a, e = e0 + tx + 16 * b;
      if (d < D && e < D) outer[(long long)d * D + e] += acc[a][b];
    }
  if (blockIdx.y == 0 && threadIdx.x < 32 && e0 + threadIdx.x < D) sum[e0 + threadIdx.x] += colsum;
}

// --- condition construction -----------------------------------------------------------------------
// mode 0 (inpaint): cond = images, box := pad.  mode 1 (outpaint): cond = pad, box := images.
__global__ void box_condition_kernel(float* __restrict__ cond, const float* __restrict__ images,
                                     const int* __restrict__ boxes, int B, int C, int H, int W, int patch,
                                     float pad_value, int mode) {
  const long long n = (long long)B * C * H * W;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int b = (int)(i / ((long long)C * H * W));
    const int h0 = boxes[2 * b], w0 = boxes[2 * b + 1];
    const bool inside = (y >= h0 && y < h0 + patch && x >= w0 && x < w0 + patch);
    cond[i] = (inside != (mode == 1)) ? pad_value : images[i];
  }
}

}  // namespace cfm
