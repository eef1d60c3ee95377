/*
 * cfm_b200.h - C ABI of the B200-native sampling engine (libcfm_b200.so).
 *
 * Drop-in boundary for ONE hot path of the reference: the iterative generation loop
 * that calls the conditional U-Net once per function evaluation (NFE).  Plain pointers
 * and sizes only; no torch types.  All tensor pointers named *_dev are CUDA device
 * pointers borrowed for the duration of the call; `stream` is a cudaStream_t passed as
 * void* (NULL = legacy default stream).  Every entry point returns 0 on success or a
 * negative cfm_status; the message is available from cfm_last_error().  Nothing throws
 * across this boundary.  One engine per GPU; an engine is not thread-safe.
 *
 * Reference interfaces replaced (paths relative to the reference repo root,
 * AD = "amortised diffusion"):
 *
 *   cfm_engine_create      <- UNetModel.__init__ / create_model / UNetModelWrapper(...)
 *                             AD/image_diffusion/unet.py:43-125, 518-706;
 *                             cifar10/compute_fid.py:39-64 (constructor + load_state_dict)
 *   cfm_engine_forward     <- UNetModel.forward(x, timesteps)   AD/image_diffusion/unet.py:708-728
 *                             UNetModelWrapper.forward(t, x, y) cifar10/compute_fid.py:70,83
 *                             model.forward(x, t, con=)         mnist/utils_mnist.py:97
 *                             eps_model(xi, i)                  AD/experiments/main.py:140
 *   cfm_sample_euler       <- NeuralODE(model,"euler").trajectory(x, t_span)
 *                             cifar10/compute_fid.py:78-79, cifar10/utils_cifar.py:34-39,
 *                             mnist/utils_mnist2.py:118-134; uint8 conversion compute_fid.py:86-87
 *   cfm_sample_ddpm        <- get_prior_sample_fn / get_conditional_sample_fn(...)(xT[, condition])
 *                             AD/image_diffusion/sampling.py:50-75, 80-133, 209-260
 *                             with DDPM tables AD/image_diffusion/sde_diffusion.py:127-167
 *   cfm_rk_combine, cfm_rk_error_sumsq, cfm_rk_scaled_sumsq, cfm_rk_dense_output
 *                          <- the state algebra of torchdiffeq.odeint(method="dopri5")
 *                             cifar10/compute_fid.py:83-85, mnist/utils_mnist.py:101-108
 *                             (the accept/reject controller stays on the host)
 *   cfm_make_box_condition <- InPainting/OutPainting._sample  AD/image_diffusion/likelihoods.py:78-105
 *   cfm_ddpm_step          <- one iteration of the reverse chains above around ANY `eps_model(xi, i)` callable
 *                             (type Network, AD/image_diffusion/sde_diffusion.py:11; AD/experiments/main.py:140)
 *   cfm_sample_sde, cfm_sde_em_step
 *                          <- torchsde.sdeint(SDE(model, score_model), x0, ts, dt=0.01)  conditional_mnist.ipynb cells 11-12
 *   cfm_ddpm_em_step       <- em_step                        AD/image_diffusion/sampling.py:100-111
 *   cfm_resize_bilinear    <- HyperResolution._sample / downsample_images
 *                             AD/image_diffusion/likelihoods.py:119-126, mnist/utils_mnist_hy.py:18-28
 *   cfm_fid_accumulate     <- the mu / Sigma accumulation behind fid.compute_fid / FrechetInceptionDistance.update
 *                             cifar10/compute_fid.py:92-100, AD/experiments/main.py:261-267, 292-293
 */
#ifndef CFM_B200_H_
#define CFM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFM_ABI_VERSION 2
#define CFM_MAX_LEVELS 8

typedef enum cfm_status {
  CFM_OK = 0,
  CFM_ERR_INVALID = -1,   /* bad argument / unsupported configuration            */
  CFM_ERR_MISSING = -2,   /* a state_dict tensor is missing or has the wrong size */
  CFM_ERR_CUDA = -3,      /* CUDA runtime / driver error                          */
  CFM_ERR_OOM = -4,
  CFM_ERR_INTERNAL = -5
} cfm_status;

typedef enum cfm_precision {
  CFM_PRECISION_FP32 = 0, /* fp32 storage + fp32 CUDA-core math: <=1e-4 rel-L2 per NFE     */
  CFM_PRECISION_BF16 = 1  /* bf16 storage, tcgen05 tensor cores, fp32 accumulate: <=2e-2   */
} cfm_precision;

/* Hyper-parameters of UNetModel.__init__ (AD/image_diffusion/unet.py:518-539). */
typedef struct cfm_unet_config {
  int32_t image_size;
  int32_t in_channels;          /* total U-Net input channels (x channels + conditioning channels) */
  int32_t model_channels;
  int32_t out_channels;
  int32_t num_res_blocks;
  int32_t n_levels;
  float   channel_mult[CFM_MAX_LEVELS];
  int32_t n_attention_ds;
  int32_t attention_ds[CFM_MAX_LEVELS];
  int32_t conv_resample;
  int32_t num_classes;          /* 0 = not class conditional */
  int32_t num_heads;
  int32_t num_head_channels;    /* -1 = use num_heads */
  int32_t num_heads_upsample;   /* -1 = num_heads */
  int32_t use_scale_shift_norm;
  int32_t resblock_updown;
  int32_t use_new_attention_order;
  int32_t precision;            /* cfm_precision */
  int32_t flags;                /* CFM_FLAG_* (0 = defaults) */
  int32_t reserved[6];
} cfm_unet_config;

/* keep every GroupNorm a pass of its own (by default a ResBlock's out_layers GroupNorm + SiLU is applied in the epilogue
 * of the block's first conv whenever the tensor-core kernel can hold the statistics of whole samples) */
#define CFM_FLAG_SEPARATE_GROUPNORM 1

typedef struct cfm_engine cfm_engine;

/* Message for the most recent failure on `e` (or, with e == NULL, of the last failed
 * cfm_engine_create on this thread).  Never NULL. */
const char* cfm_last_error(const cfm_engine* e);
int cfm_abi_version(void);

/* Build an engine on CUDA device `device` from a state_dict: `names[i]` follows the
 * reference's module paths ("input_blocks.1.0.in_layers.2.weight", ...), `host_data[i]`
 * points at `numel[i]` contiguous fp32 values in HOST memory (row-major, PyTorch layout).
 * The engine copies and repacks; the caller keeps ownership of the inputs. */
int cfm_engine_create(const cfm_unet_config* cfg, int32_t n_tensors, const char* const* names,
                      const float* const* host_data, const int64_t* numel, int32_t device,
                      cfm_engine** out);
void cfm_engine_destroy(cfm_engine* e);

/* Introspection (for roofline accounting and tests). */
int64_t cfm_engine_param_count(const cfm_engine* e);
double  cfm_engine_flops_per_sample(const cfm_engine* e);       /* 2*MAC, conv+linear+attention  */
int64_t cfm_engine_workspace_bytes(const cfm_engine* e, int32_t batch);
int32_t cfm_engine_kernel_launches(const cfm_engine* e);        /* launches issued by the last call */
int32_t cfm_engine_tensor_core_convs(const cfm_engine* e);      /* conv ops routed to tcgen05 per NFE */
int32_t cfm_engine_cached_graphs(const cfm_engine* e);          /* captured step graphs held (LRU-bounded)  */

/* One NFE.  x_dev: [B, Cx, H, W] fp32 NCHW.  cond_dev: [B, Cc, H, W] fp32 NCHW or NULL
 * (Cx + Cc == in_channels; cond is concatenated after x on the channel axis, as the
 * reference's InPaint/Amortized wrappers do).  Time: if t_dev != NULL it is a device
 * array of B per-sample fp32 timesteps, else `t_scalar` is used for every sample
 * (UNetModelWrapper's 0-dim t).  y_dev: B int64 class labels or NULL (must be non-NULL
 * iff num_classes > 0).  out_dev: [B, out_channels, H, W] fp32 NCHW. */
int cfm_engine_forward(cfm_engine* e, int32_t batch, const float* x_dev, const float* cond_dev,
                       const float* t_dev, float t_scalar, const int64_t* y_dev,
                       float* out_dev, void* stream);

/* Per-op device timing of one NFE (CUDA events around every launch, averaged over `repeats`
 * evaluations after one warm-up).  Results are read back with cfm_engine_profile_count/_get;
 * kind: 0 generic conv, 1 groupnorm, 2 resample, 3 generic attention, 4 tcgen05 conv, 5 tcgen05 attention. */
int cfm_engine_profile_forward(cfm_engine* e, int32_t batch, const float* x_dev, const float* cond_dev,
                               float t_scalar, const int64_t* y_dev, float* out_dev, int32_t repeats,
                               void* stream);
int32_t cfm_engine_profile_count(const cfm_engine* e);
int cfm_engine_profile_get(const cfm_engine* e, int32_t i, char* name, int32_t name_cap, int32_t* kind,
                           double* ms, double* flops_per_sample);
/* Static accounting of op i of the plan (same index space as cfm_engine_profile_get), per sample:
 *   what = 0: EXECUTED flops (2*MAC the kernel issues: 4/9 of the algorithmic count for the folded
 *             "nearest x2 upsample + 3x3 conv", more than it for zero-padded K / N of the stem and head GEMMs)
 *   what = 1: algorithmic bytes (each operand read once + the result written once, at the storage width) */
int cfm_engine_op_info(const cfm_engine* e, int32_t i, int32_t what, double* value);

/* Flags for cfm_sample_euler. */
#define CFM_EULER_COND_DRIFT   1u  /* conditioning is ODE state with d(con)/dt = con */
#define CFM_EULER_USE_GRAPH    2u  /* capture ONE step (U-Net + update + counter bump) as a CUDA graph and replay it
                                    * n_steps times; per-step scalars come from device tables indexed by a step counter */

/* Fixed-step Euler: for k in [0, n_steps): x += dt[k] * model(t[k], x, y, cond).
 * t_host / dt_host: n_steps fp32 values each, in HOST memory (the torchdyn grid).
 * x_dev is updated IN PLACE and holds the final state on return.
 * traj_dev (optional): [n_steps + 1, B, Cx, H, W] fp32, receives every state incl. x0.
 * img_u8_dev (optional): [B, Cx, H, W] uint8 = clip(x*127.5 + 128, 0, 255) of the final state. */
int cfm_sample_euler(cfm_engine* e, int32_t batch, float* x_dev, float* cond_dev,
                     const int64_t* y_dev, const float* t_host, const float* dt_host,
                     int32_t n_steps, uint32_t flags, float* traj_dev, uint8_t* img_u8_dev,
                     void* stream);

/* Classifier-free guidance on the same loop (BASELINE config 3; an EXTENSION - the reference's conditional_mnist
 * notebook evaluates model(t, x, y) only): two U-Net evaluations per step,
 *   v_c = model(t, x, y),  v_u = model(t, x) with the label embedding left out,  x += dt * (v_c + w (v_c - v_u)).
 * The model must be class-conditional; guidance_w = 0 reproduces cfm_sample_euler. */
int cfm_sample_euler_cfg(cfm_engine* e, int32_t batch, float* x_dev, float* cond_dev,
                         const int64_t* y_dev, float guidance_w, const float* t_host, const float* dt_host,
                         int32_t n_steps, uint32_t flags, float* traj_dev, uint8_t* img_u8_dev,
                         void* stream);

typedef enum cfm_ddpm_mode {
  CFM_DDPM_PRIOR = 0,        /* sampling.py:50-75   */
  CFM_DDPM_REPLACEMENT = 1,  /* sampling.py:209-260 */
  CFM_DDPM_AMORTIZED = 2     /* sampling.py:80-133  */
} cfm_ddpm_mode;

/* Per-step scalars, each an Ns-long fp32 HOST array (the DDPM buffers of
 * sde_diffusion.py:127-167) plus the model time of each step (main.py:140). */
typedef struct cfm_ddpm_tables {
  int32_t Ns;
  const float* sqrt_alphas_cumprod;
  const float* sqrt_one_minus_alphas_cumprod;
  const float* sqrt_recip_alphas_cumprod;
  const float* sqrt_recipm1_alphas_cumprod;
  const float* posterior_mean_coef1;
  const float* posterior_mean_coef2;
  const float* posterior_log_variance_clipped;
  const float* model_time;           /* t fed to the U-Net at step i (i / Ns) */
} cfm_ddpm_tables;

typedef struct cfm_ddpm_options {
  int32_t mode;                /* cfm_ddpm_mode */
  float   pad_value;           /* Replacement: mask marker, exact compare (likelihoods.py pad_value = -2);
                                * Amortized with correctors: the constant of likelihood.none_like() (sampling.py:36-37) */
  int32_t replace_below_step;  /* blend while i < this (int(Ns * start_fraction))            */
  int32_t noise_condition;     /* Replacement: q_sample the condition (1) or use it raw (0)  */
  uint32_t use_graph;
  uint32_t n_corrector;        /* Langevin corrector steps after every predictor step (Replacement: sampling.py:241-256;
                                * Amortized: sampling.py:113-127, the corrector's U-Net call sees none_like() as condition) */
  float   corrector_delta;     /* conditioning.delta: step = 0.5*dt*delta*score + sqrt(dt*delta)*z, dt = (1 - 1e-5)/Ns      */
  uint32_t reserved[1];
} cfm_ddpm_options;

/* Reverse chain i = Ns-1 .. 0.  x_dev [B,C,H,W] holds xT on entry and clip(x0,-1,1) on
 * return.  condition_dev: [B,C,H,W] (Replacement: image with pad_value holes; Amortized:
 * the conditioning image, concatenated on channels) or NULL for PRIOR.
 * Noise: if noise_dev != NULL it is [Ns, 2 + n_corrector, B*C*H*W] fp32: slot (i,0) is the q_sample
 * draw for the mask blend at step i, slot (i,1) the posterior draw at step i, slots (i,2..) the
 * corrector draws (the reference's randn_like calls, in call order).  Otherwise a counter-based Philox
 * generator seeded with `seed` is used on the device. */
int cfm_sample_ddpm(cfm_engine* e, int32_t batch, float* x_dev, const float* condition_dev,
                    const cfm_ddpm_tables* tables, const cfm_ddpm_options* opt,
                    const float* noise_dev, uint64_t seed, void* stream);

/* dopri5 state algebra on the device (host keeps the controller).
 * out = y + dt * sum_j coef[j] * k[j]   over n fp32 elements; k_dev: HOST array of n_k device pointers. */
int cfm_rk_combine(float* out_dev, const float* y_dev, const float* const* k_dev,
                   const float* coef_host, int32_t n_k, float dt, int64_t n, void* stream);
/* sumsq_dev[0] (double, device) = sum_i ( err_i / (atol + rtol*max(|y0_i|,|y1_i|)) )^2, err = dt*sum_j coef[j]*k[j]. */
int cfm_rk_error_sumsq(double* sumsq_dev, const float* y0_dev, const float* y1_dev,
                       const float* const* k_dev, const float* coef_host, int32_t n_k, float dt,
                       float rtol, float atol, int64_t n, void* stream);

/* dopri5 initial-step heuristic (torchdiffeq `_select_initial_step`, called from odeint: compute_fid.py:83-85,
 * utils_mnist.py:101-108): sumsq_dev[0] (double, device) = sum_i ((a_i - b_i) / (atol + rtol*|y_i|))^2; b_dev may be NULL. */
int cfm_rk_scaled_sumsq(double* sumsq_dev, const float* a_dev, const float* b_dev, const float* y_dev,
                        float rtol, float atol, int64_t n, void* stream);
/* dopri5 dense output (torchdiffeq `_interp_fit` / `_interp_evaluate`): the quartic through y0, y_mid, y1 with end
 * slopes f0, f1 of an accepted step of length dt, evaluated at fraction x in [0, 1] of the step. */
int cfm_rk_dense_output(float* out_dev, const float* y0_dev, const float* y1_dev, const float* ymid_dev,
                        const float* f0_dev, const float* f1_dev, float dt, float x, int64_t n, void* stream);

/* Box-mask condition on the device: boxes_dev[b] = {h, w} (int32 pairs, drawn on the host
 * with the reference's RNG order).  inpaint: cond = images with box := pad_value;
 * outpaint (mode 1): cond = pad_value everywhere except the box. */
int cfm_make_box_condition(float* cond_dev, const float* images_dev, const int32_t* boxes_dev,
                           int32_t batch, int32_t channels, int32_t height, int32_t width,
                           int32_t patch, float pad_value, int32_t mode, void* stream);

/* clip(x*127.5+128, 0, 255) -> uint8 (compute_fid.py:87). */
int cfm_quantize_u8(uint8_t* out_dev, const float* x_dev, int64_t n, void* stream);

/* ONE launch of a DDPM reverse-chain step, for a caller that evaluates the eps network itself between the launches -
 * the reference's Network seam is any callable `eps_model(xi, i)` (sde_diffusion.py:11; main.py:140 passes a lambda).
 * Same tables / options / noise layout / Philox streams as cfm_sample_ddpm, so a chain driven through this entry point
 * with the engine's own forward as the network equals cfm_sample_ddpm bit for bit.
 *   phase 0      mask blend of step `chain_index` (Replacement, before the network call; no-op otherwise)
 *   phase 1      posterior draw: x <- c1*clip(a x - b eps) + c2 x + sigma z           (sampling.py:59-67)
 *   phase 2 + c  Langevin corrector c: eps is the network re-evaluated on the current x   (sampling.py:113-121, 241-250)
 * With opt->n_corrector == 0 the phase-1 launch of step 0 applies the final clip, else the last corrector of step 0. */
int cfm_ddpm_step(float* x_dev, const float* eps_dev, const float* condition_dev, const cfm_ddpm_tables* tables,
                  const cfm_ddpm_options* opt, int32_t chain_index, int32_t phase, const float* noise_dev,
                  uint64_t seed, int64_t n, void* stream);

/* Euler-Maruyama steps.
 * cfm_sde_em_step:  x <- x + (drift + score) dt + sigma sqrt(dt) z   - the "euler" scheme torchsde.sdeint applies to
 *   SDE.f = flow + score, SDE.g = sigma (conditional_mnist.ipynb cells 11-12); score_dev may be NULL.
 * cfm_ddpm_em_step: the reverse VP-SDE step `em_step` of the Amortized sampler (sampling.py:100-111 with
 *   sde_diffusion.py:170-205): score = -eps / sigma_t, drift = -0.5 x x - beta_t score (the reference's DDPM.drift
 *   swaps the arguments of unsqueeze_like and so multiplies by x instead of beta_t; reproduced as is),
 *   x <- x - dt drift + sqrt(beta_t) z sqrt(dt).
 * z: noise_dev (n injected normals) or Philox(seed, stream_id). */
int cfm_sde_em_step(float* x_dev, const float* drift_dev, const float* score_dev, float dt, float sigma,
                    const float* noise_dev, uint64_t seed, uint32_t stream_id, int64_t n, void* stream);
int cfm_ddpm_em_step(float* x_dev, const float* eps_dev, float beta_t, float sigma_t, double dt, const float* noise_dev,
                     uint64_t seed, uint32_t stream_id, int64_t n, void* stream);
/* Whole fixed-step Euler-Maruyama loop over two engines (flow + score networks; `score` may be NULL): for k in
 * [0, n_steps): x += (drift(t[k], x, y) + score(t[k], x, y)) dt[k] + sigma sqrt(dt[k]) z_k.  x_dev in place.
 * noise_dev: [n_steps, B*C*H*W] or NULL (Philox(seed), stream = k). */
int cfm_sample_sde(cfm_engine* drift, cfm_engine* score, int32_t batch, float* x_dev, const int64_t* y_dev,
                   const float* t_host, const float* dt_host, int32_t n_steps, float sigma, const float* noise_dev,
                   uint64_t seed, void* stream);

/* F.interpolate(mode="bilinear", align_corners=False) on `planes` = B*C fp32 planes of h_in x w_in -> h_out x w_out:
 * HyperResolution._sample (likelihoods.py:119-126), downsample_images (mnist/utils_mnist_hy.py:18-28) and the
 * low-res -> full-size upsample of SuperResModelWrapper. */
int cfm_resize_bilinear(float* out_dev, const float* in_dev, int64_t planes, int32_t h_in, int32_t w_in,
                        int32_t h_out, int32_t w_out, void* stream);

/* FID sufficient statistics (cifar10/compute_fid.py:92-100; AD/experiments/main.py:261-267, 292-293): running fp64
 * sums over feature rows, sum_dev[dim] += sum_i f_i, outer_dev[dim][dim] += sum_i f_i f_i^T; feats_dev: [n, dim] fp32.
 * Deterministic (one accumulating thread per output element, row order).  The sums are what ranks all-reduce. */
int cfm_fid_accumulate(double* sum_dev, double* outer_dev, const float* feats_dev, int64_t n, int32_t dim, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CFM_B200_H_ */
