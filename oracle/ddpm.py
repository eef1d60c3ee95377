"""CPU oracle: DDPM tables, posterior algebra, reverse-chain samplers, condition makers.

TEST INFRASTRUCTURE ONLY.  Follows (AD = /root/reference/amortised diffusion):

* tables ...................... AD/image_diffusion/sde_diffusion.py:127-167 (DDPM.__init__), beta :25-27
* extract ..................... AD/image_diffusion/sde_diffusion.py:101-104
* predict_start / posterior ... AD/image_diffusion/sde_diffusion.py:220-237
* q_sample .................... AD/image_diffusion/sde_diffusion.py:239-244
* prior sampler ............... AD/image_diffusion/sampling.py:50-75
* Amortized sampler ........... AD/image_diffusion/sampling.py:30-44, 80-133 (predictor + Langevin corrector)
* Replacement sampler ......... AD/image_diffusion/sampling.py:209-260 (mask blend at :232)
* eps wrapper t = 1.0*i/Ns .... AD/experiments/main.py:140
* InPainting / OutPainting .... AD/image_diffusion/likelihoods.py:39-105 ; twins mnist/utils_mnist.py:16-41
* HyperResolution ............. AD/image_diffusion/likelihoods.py:108-126 ; mnist/utils_mnist_hy.py:18-28

``sampling.py`` itself cannot be imported (needs ``plum``), so the samplers are
pinned only through ``sde_diffusion.DDPM`` (importable: tables and step algebra are
compared bit-for-bit in tests/test_oracle_vs_reference.py); the loop order around
them is restated.  Random draws are INJECTED (``noise(shape)`` is called exactly
where the reference calls ``torch.randn_like``, in the same order) so an engine
fed the same tensors must agree.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F

BM, BD = 0.1, 20


def ddpm_tables(Ns: int) -> Dict[str, torch.Tensor]:
    ts = torch.linspace(0.00001, 1.0, Ns, dtype=torch.float32)
    betas = ((BM + (BD - BM) * ts) / Ns).to(torch.float32)
    alphas = (1.0 - betas).to(torch.float32)
    ac = torch.cumprod(alphas, dim=0)
    ac_prev = F.pad(ac[:-1], (1, 0), value=1.0)
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    return {
        "ts": ts, "alphas": alphas, "betas": betas,
        "alphas_cumprod": ac, "alphas_cumprod_prev": ac_prev,
        "sqrt_alphas_cumprod": torch.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - ac),
        "log_one_minus_alphas_cumprod": torch.log(1.0 - ac),
        "sqrt_recip_alphas_cumprod": torch.sqrt(1.0 / ac),
        "sqrt_recipm1_alphas_cumprod": torch.sqrt(1.0 / ac - 1),
        "recip_sqrt_m1_alphas_cumprod": 1.0 / torch.sqrt(1 - ac),
        "posterior_variance": post_var,
        "posterior_log_variance_clipped": torch.log(post_var.clamp(min=1e-20)),
        "posterior_mean_coef1": betas * torch.sqrt(ac_prev) / (1.0 - ac),
        "posterior_mean_coef2": (1.0 - ac_prev) * torch.sqrt(alphas) / (1.0 - ac),
    }


def eps_time(i: int, Ns: int) -> torch.Tensor:
    """main.py:140 - ``1.0 * i / Ns`` with ``i`` an int64 tensor: fp32 multiply then fp32 divide."""
    return (1.0 * torch.tensor(i, dtype=torch.long)) / Ns


def posterior_step(tb, xi, eps, i: int, z):
    """One predictor step given eps: x0 = clip(a*xi - b*eps); mean = c1*x0 + c2*xi; + sigma*z."""
    x0 = torch.clip(tb["sqrt_recip_alphas_cumprod"][i] * xi - tb["sqrt_recipm1_alphas_cumprod"][i] * eps, -1, 1)
    mean = tb["posterior_mean_coef1"][i] * x0 + tb["posterior_mean_coef2"][i] * xi
    scale = (0.5 * tb["posterior_log_variance_clipped"][i]).exp()
    return mean + scale * z, x0


def _eps_call(eps_model, xi, i, Ns):
    return eps_model(xi, eps_time(i, Ns).repeat(xi.shape[0]))


@torch.no_grad()
def sample_prior(eps_model: Callable, Ns: int, xT: torch.Tensor, noise: Callable) -> torch.Tensor:
    tb = ddpm_tables(Ns)
    xi = xT
    for i in reversed(range(Ns)):
        eps = _eps_call(eps_model, xi, i, Ns)
        z = noise(xi.shape) if i > 0 else 0.0
        xi, _ = posterior_step(tb, xi, eps, i, z)
    return torch.clip(xi, -1, 1)


@torch.no_grad()
def sample_replacement(eps_model: Callable, Ns: int, xT, condition, noise: Callable, pad_value: float = -2.0,
                       start_fraction: float = 1.0, noise_condition: bool = True,
                       n_corrector: int = 0, delta: float = 0.1) -> torch.Tensor:
    tb = ddpm_tables(Ns)
    xi = xT
    for i in reversed(range(Ns)):
        if i < int(Ns * start_fraction):
            if noise_condition:
                z1 = noise(condition.shape)        # q_sample's randn_like - drawn FIRST
                nc = tb["sqrt_alphas_cumprod"][i] * condition + tb["sqrt_one_minus_alphas_cumprod"][i] * z1
            else:
                nc = condition
            xi = torch.where(condition == pad_value, xi, nc)
        eps = _eps_call(eps_model, xi, i, Ns)
        z = noise(xi.shape) if i > 0 else 0.0
        xi, _ = posterior_step(tb, xi, eps, i, z)
        for _ in range(n_corrector):
            xi = corrector_step(tb, eps_model, Ns, xi, i, delta, noise)
    return torch.clip(xi, -1, 1)


@torch.no_grad()
def sample_amortized(eps_model: Callable, Ns: int, xT, condition, noise: Callable,
                     n_corrector: int = 0, delta: float = 0.1) -> torch.Tensor:
    tb = ddpm_tables(Ns)
    xi = xT
    cat_model = lambda x, t: eps_model(torch.cat([x, condition], dim=-3), t)
    for i in reversed(range(Ns)):
        eps = _eps_call(cat_model, xi, i, Ns)
        z = noise(xi.shape) if i > 0 else 0.0
        xi, _ = posterior_step(tb, xi, eps, i, z)
        for _ in range(n_corrector):
            xi = corrector_step(tb, cat_model, Ns, xi, i, delta, noise)
    return torch.clip(xi, -1, 1)


def corrector_step(tb, eps_model, Ns, xi, i, delta, noise):
    """Langevin corrector (sampling.py:113-121): score = -x0_hat / sqrt(1-abar) (sde_diffusion.py:214-217)."""
    eps = _eps_call(eps_model, xi, i, Ns)
    x0 = torch.clip(tb["sqrt_recip_alphas_cumprod"][i] * xi - tb["sqrt_recipm1_alphas_cumprod"][i] * eps, -1, 1)
    score = -tb["recip_sqrt_m1_alphas_cumprod"][i] * x0
    dt = (1.0 - 0.00001) / Ns
    drift = 0.5 * dt * delta * score
    nz = math.sqrt(dt * delta) * noise(xi.shape)
    return xi + (drift + nz)            # the reference writes ``xi += drift + noise``


# --- condition construction --------------------------------------------------------------

def random_patch(image_size: int, patch_size: int, generator: Optional[torch.Generator] = None):
    """likelihoods.py:49-53 - h first, then w, both ``randint(5, S - p - 5)`` on the CPU generator."""
    h = torch.randint(5, image_size - patch_size - 5, size=(), generator=generator)
    w = torch.randint(5, image_size - patch_size - 5, size=(), generator=generator)
    return int(h), int(w)


def inpainting_condition(images, patch_size: int, pad_value: float = -2.0, generator=None, boxes=None):
    """Per-sample box := pad_value (likelihoods.py:78-87 applied sample by sample, :22-27)."""
    cond = images.detach().clone()
    S = images.shape[-1]
    for k in range(images.shape[0]):
        h, w = random_patch(S, patch_size, generator)
        if boxes is not None:
            boxes.append((h, w))
        cond[k, :, h:h + patch_size, w:w + patch_size] = pad_value
    return cond


def outpainting_condition(images, patch_size: int, pad_value: float = -2.0, generator=None, boxes=None):
    """Everything masked except the box (likelihoods.py:90-105)."""
    cond = torch.ones_like(images) * pad_value
    S = images.shape[-1]
    for k in range(images.shape[0]):
        h, w = random_patch(S, patch_size, generator)
        if boxes is not None:
            boxes.append((h, w))
        cond[k, :, h:h + patch_size, w:w + patch_size] = images[k, :, h:h + patch_size, w:w + patch_size]
    return cond


def downsample_images(images, target_size):
    """mnist/utils_mnist_hy.py:18-28."""
    return F.interpolate(images, size=target_size, mode="bilinear", align_corners=False)


def hyperresolution_condition(images, target_hw):
    """likelihoods.py:119-126: bilinear down, then bilinear back up to full size."""
    lo = F.interpolate(images, size=tuple(target_hw), mode="bilinear", align_corners=False)
    return F.interpolate(lo, (images.shape[2], images.shape[3]), mode="bilinear")


def to_uint8(x):
    """cifar10/compute_fid.py:87."""
    return (x * 127.5 + 128).clip(0, 255).to(torch.uint8)
