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

// --- Euler:  x <- x + dt * v   (torchdyn fixed-step driver) ----------------------------------------
// cond (optional, COND_DRIFT): con <- con + dt * con  (the reference's ode_func returns x[1] as d(con)/dt)
// traj (optional): next trajectory slot receives the new x.  img (optional): uint8 quantisation.
__global__ void euler_step_kernel(float* __restrict__ x, const float* __restrict__ v, const float* __restrict__ dt_table,
                                  const int* __restrict__ step_counter, int n_steps, long long n,
                                  float* __restrict__ cond, long long n_cond,
                                  float* __restrict__ traj, uint8_t* __restrict__ img) {
  const int k = *step_counter;
  const float dt = dt_table[k];
  float* traj_slot = traj ? traj + (long long)(k + 1) * n : nullptr;
  const bool want_img = img && k == n_steps - 1;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float nx = __fadd_rn(x[i], __fmul_rn(dt, v[i]));
    x[i] = nx;
    if (traj_slot) traj_slot[i] = nx;
    if (want_img) img[i] = (uint8_t)fminf(fmaxf(__fadd_rn(__fmul_rn(nx, 127.5f), 128.0f), 0.0f), 255.0f);
  }
  if (cond)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_cond; i += stride)
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
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = (uint8_t)fminf(fmaxf(__fadd_rn(__fmul_rn(x[i], 127.5f), 128.0f), 0.0f), 255.0f);
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
// The step's scalars come from a device table indexed by the device-side step counter, so one captured
// launch serves every step.  noise (optional): [Ns, 2, n]; slot (i,0) = blend draw, (i,1) = posterior draw;
// without it a Philox generator is used (streams 2i / 2i+1).
__global__ void ddpm_step_kernel(float* __restrict__ x, const float* __restrict__ eps,
                                 const DdpmStepScalars* __restrict__ table, const int* __restrict__ step_counter,
                                 const float* __restrict__ cond, const float* __restrict__ noise,
                                 unsigned long long seed, long long n) {
  const DdpmStepScalars s = table[*step_counter];
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

// Blend of the CURRENT step from the device table (chains with corrector steps): x = where(cond == pad, x, q_sample(cond)).
__global__ void ddpm_blend_table_kernel(float* __restrict__ x, const float* __restrict__ cond,
                                        const DdpmStepScalars* __restrict__ table, const int* __restrict__ step_counter,
                                        const float* __restrict__ noise, unsigned long long seed, long long n) {
  const DdpmStepScalars s = table[*step_counter];
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
__global__ void ddpm_corrector_kernel(float* __restrict__ x, const float* __restrict__ eps,
                                      const DdpmStepScalars* __restrict__ table, const int* __restrict__ step_counter,
                                      const float* __restrict__ noise, unsigned long long seed, int slot, int last, long long n) {
  const DdpmStepScalars s = table[*step_counter];
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

__global__ void rk_error_sumsq_kernel(double* __restrict__ sumsq, const float* __restrict__ y0, const float* __restrict__ y1,
                                      RkPtrs p, float dt, float rtol, float atol, long long n) {
  __shared__ double red[32];
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
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = acc;
  __syncthreads();
  if (w == 0) {
    acc = (lane < (int)(blockDim.x >> 5)) ? red[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) atomicAdd(sumsq, acc);
  }
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
