"""DDPM reverse-chain sampling over the engine.

Mirrors the reference's operator surface for this path:

* ``DDPM(Ns)`` buffers and algebra ........ AD/image_diffusion/sde_diffusion.py:107-244
* ``Amortized`` / ``Replacement`` ......... AD/image_diffusion/conditioning.py:12-63
* ``InPainting`` / ``OutPainting`` / ``HyperResolution`` ... AD/image_diffusion/likelihoods.py:39-146
* ``get_prior_sample_fn`` / ``get_conditional_sample_fn(eps_model, ddpm, conditioning, likelihood)``
  ........................................ AD/image_diffusion/sampling.py:50-75, 80-133, 209-260
* ``EpsModel(network, ddpm)`` is the object form of ``lambda xi, i: network(xi, 1.0 * i / ddpm.Ns)``
  (AD/experiments/main.py:140) that lets the sampler see the engine and run the whole chain
  in one native call.  ``ReconstructionGuidance`` needs the U-Net's backward pass and is out
  of scope for this inference engine (raises NotImplementedError).

The multiple-dispatch on (conditioning, likelihood) types that the reference does with ``plum``
is done with ``isinstance`` here.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Callable, Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from .models import UNetModel

bm = 0.1
bd = 20


def beta(t):
    return bm + (bd - bm) * t


def extract(a, t, x_shape):
    b, *_ = t.shape
    out = a.gather(-1, t)
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))


class DDPM(nn.Module):
    """Time-discretised VP SDE tables; buffer names match the reference so checkpoints/configs carry over."""

    def __init__(self, Ns: int):
        super().__init__()
        self.Ns = Ns
        self.tmin, self.tmax = 0.00001, 1.0
        self.ts = torch.linspace(self.tmin, self.tmax, Ns, dtype=torch.float32)
        reg = lambda name, val: self.register_buffer(name, val.to(torch.float32))
        betas = beta(self.ts) / Ns
        reg("alphas", 1.0 - betas)
        ac = torch.cumprod(self.alphas, dim=0)
        ac_prev = F.pad(ac[:-1], (1, 0), value=1.0)
        reg("betas", betas); reg("alphas_cumprod", ac); reg("alphas_cumprod_prev", ac_prev)
        reg("sqrt_alphas_cumprod", torch.sqrt(ac))
        reg("sqrt_one_minus_alphas_cumprod", torch.sqrt(1.0 - ac))
        reg("log_one_minus_alphas_cumprod", torch.log(1.0 - ac))
        reg("sqrt_recip_alphas_cumprod", torch.sqrt(1.0 / ac))
        reg("sqrt_recipm1_alphas_cumprod", torch.sqrt(1.0 / ac - 1))
        reg("recip_sqrt_m1_alphas_cumprod", 1.0 / torch.sqrt(1 - ac))
        pv = betas * (1.0 - ac_prev) / (1.0 - ac)
        reg("posterior_variance", pv)
        reg("posterior_log_variance_clipped", torch.log(pv.clamp(min=1e-20)))
        reg("posterior_mean_coef1", betas * torch.sqrt(ac_prev) / (1.0 - ac))
        reg("posterior_mean_coef2", (1.0 - ac_prev) * torch.sqrt(self.alphas) / (1.0 - ac))

    def model_time(self) -> torch.Tensor:
        """t fed to the U-Net at step i: ``1.0 * i / Ns`` on an int64 tensor (main.py:140)."""
        return (1.0 * torch.arange(self.Ns, dtype=torch.long)) / self.Ns

    def tables(self) -> dict:
        names = ("sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
                 "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2",
                 "posterior_log_variance_clipped")
        tb = {n: getattr(self, n).detach().cpu() for n in names}
        tb["model_time"] = self.model_time()
        return tb

    # algebra kept for API compatibility (host/torch tensors; the fused kernel is used by the samplers)
    def predict_start_from_noise(self, x_i, i, noise):
        return extract(self.sqrt_recip_alphas_cumprod, i, x_i.shape) * x_i - extract(self.sqrt_recipm1_alphas_cumprod, i, x_i.shape) * noise

    def q_posterior(self, x0, x_i, i):
        mean = extract(self.posterior_mean_coef1, i, x_i.shape) * x0 + extract(self.posterior_mean_coef2, i, x_i.shape) * x_i
        return mean, extract(self.posterior_variance, i, x_i.shape), extract(self.posterior_log_variance_clipped, i, x_i.shape)

    def p_mean_variance(self, x_start, x, i):
        m, v, lv = self.q_posterior(x0=x_start, x_i=x, i=i)
        return m, v, lv, x_start

    def q_sample(self, x_start, i):
        noise = torch.randn_like(x_start)
        return extract(self.sqrt_alphas_cumprod, i, x_start.shape) * x_start + extract(self.sqrt_one_minus_alphas_cumprod, i, x_start.shape) * noise, noise


# --- conditioning strategies (conditioning.py) ----------------------------------------------------
class Conditioning:
    @classmethod
    def from_configdict(cls, config):
        return cls()


class Amortized(Conditioning):
    def __init__(self, p_cond: float = 0.9, n_corrector: int = 0, delta: float = 0.1):
        self.p_cond, self.n_corrector, self.delta = p_cond, n_corrector, delta

    @classmethod
    def from_configdict(cls, config):
        return cls(p_cond=config["p_cond"], n_corrector=config["n_corrector"], delta=config["delta"])


class Replacement(Conditioning):
    def __init__(self, delta: float = 0.1, start_fraction: float = 1.0, noise: bool = True, n_corrector: int = 0):
        self.delta, self.start_fraction, self.noise, self.n_corrector = delta, start_fraction, noise, n_corrector

    @classmethod
    def from_configdict(cls, config):
        return cls(delta=config["delta"], start_fraction=config["start_fraction"], noise=config["noise"],
                   n_corrector=config["n_corrector"])


class ReconstructionGuidance(Conditioning):
    def __init__(self, *a, **kw):
        raise NotImplementedError("ReconstructionGuidance differentiates through the U-Net (vmap(grad)); "
                                  "the B200 engine is inference-only")


def get_conditioning(type_: str):
    return {"amortized": Amortized, "replacement": Replacement, "reconstruction_guidance": ReconstructionGuidance}[type_.lower()]


# --- likelihoods: condition construction (likelihoods.py) ------------------------------------------
class Likelihood:
    def sample(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def none_like(self, x):
        raise NotImplementedError


class Painting(Likelihood):
    mode = 0

    @classmethod
    def from_configdict(cls, config):
        return cls(patch_size=config["patch_size"], pad_value=config["pad_value"])

    def __init__(self, patch_size: int, pad_value: float):
        self.pad_value, self.patch_size = pad_value, patch_size

    def get_random_patch(self, image_size):
        # same draws, same order, same (CPU, global) generator as the reference: h, then w
        h = torch.randint(5, image_size - self.patch_size - 5, size=())
        w = torch.randint(5, image_size - self.patch_size - 5, size=())
        return h, w

    def sample_boxes(self, batch: int, image_size: int) -> torch.Tensor:
        return torch.tensor([[int(v) for v in self.get_random_patch(image_size)] for _ in range(batch)],
                            dtype=torch.int32).reshape(batch, 2)

    def sample(self, x: torch.Tensor, boxes: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Per-sample random box; the fill is one native kernel instead of an O(B) Python loop."""
        B, Cc, H, W = x.shape
        if boxes is None:
            boxes = self.sample_boxes(B, W)
        if x.device.type != "cuda":
            raise RuntimeError("images must be on a CUDA device")
        xs = x.detach().to(torch.float32).contiguous()
        out = torch.empty_like(xs)
        bd_ = boxes.to(device=x.device, dtype=torch.int32).contiguous()
        lib = _lib.load()
        with torch.cuda.device(x.device):
            rc = lib.cfm_make_box_condition(C.c_void_p(out.data_ptr()), C.c_void_p(xs.data_ptr()), C.c_void_p(bd_.data_ptr()),
                                            B, Cc, H, W, int(self.patch_size), float(self.pad_value), self.mode,
                                            C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
        _lib.check(rc)
        return out

    def none_like(self, x):
        return torch.ones_like(x) * self.pad_value


class InPainting(Painting):
    mode = 0


class OutPainting(Painting):
    mode = 1


class HyperResolution(Likelihood):
    @classmethod
    def from_configdict(cls, config):
        return cls(config["target_height"], config["target_width"])

    def __init__(self, target_height: int, target_width: int):
        self.target_height, self.target_width = target_height, target_width

    def sample(self, images: torch.Tensor) -> torch.Tensor:
        lo = F.interpolate(images, size=(self.target_height, self.target_width), mode="bilinear", align_corners=False)
        return F.interpolate(lo, (images.shape[2], images.shape[3]), mode="bilinear")

    def none_like(self, x):
        return torch.zeros_like(x)


def get_likelihood(type_: str):
    return {"inpainting": InPainting, "outpainting": OutPainting, "hyperresolution": HyperResolution}[type_.lower()]


def downsample_images(images, target_size):
    """mnist/utils_mnist_hy.py:18-28."""
    return F.interpolate(images, size=target_size, mode="bilinear", align_corners=False)


# --- eps network + sampler factories ----------------------------------------------------------------
class EpsModel:
    """``eps_model(xi, i)`` with ``i`` int64 [B]: calls ``network(xi, 1.0 * i / Ns)`` (main.py:140)."""

    def __init__(self, network: UNetModel, ddpm: DDPM):
        self.network, self.ddpm = network, ddpm

    def __call__(self, xi, i):
        return self.network(xi, 1.0 * i / self.ddpm.Ns)


def _noise_tensor(noise, Ns, x):
    if noise is None or torch.is_tensor(noise):
        return noise
    raise TypeError("noise must be a [Ns, 2 + n_corrector, B*C*H*W] tensor or None")


def get_prior_sample_fn(eps_model, ddpm: DDPM, conditioning=None, likelihood=None, *, noise=None, seed: int = 0,
                        use_graph: bool = False) -> Callable:
    if not isinstance(eps_model, EpsModel):
        raise TypeError("wrap the network as EpsModel(network, ddpm) so the sampler can run on the engine")
    amortized = isinstance(conditioning, Amortized)

    @torch.no_grad()
    def sample(xT):
        eng = eps_model.network.engine()
        if amortized:      # x0_model substitutes likelihood.none_like(xi) for the condition (sampling.py:36-37)
            return eng.sample_ddpm(xT, ddpm.tables(), mode="amortized", condition=likelihood.none_like(xT),
                                   noise=_noise_tensor(noise, ddpm.Ns, xT), seed=seed, use_graph=use_graph)
        return eng.sample_ddpm(xT, ddpm.tables(), mode="prior", noise=_noise_tensor(noise, ddpm.Ns, xT), seed=seed,
                               use_graph=use_graph)

    return sample


def get_conditional_sample_fn(eps_model, ddpm: DDPM, conditioning, likelihood, *, noise=None, seed: int = 0,
                              use_graph: bool = False) -> Callable:
    if not isinstance(eps_model, EpsModel):
        raise TypeError("wrap the network as EpsModel(network, ddpm) so the sampler can run on the engine")
    if getattr(conditioning, "n_corrector", 0) and not isinstance(conditioning, (Replacement, Amortized)):
        raise NotImplementedError("Langevin corrector steps (n_corrector > 0) run on the engine for Replacement and "
                                  "Amortized conditioning only")

    if isinstance(conditioning, Amortized):
        n_corr = int(getattr(conditioning, "n_corrector", 0))

        @torch.no_grad()
        def sample(xT, condition):
            none_value = 0.0
            if n_corr:
                # the corrector's x0_model call has no condition -> likelihood.none_like(xi) (sampling.py:36-37, 116);
                # every reference likelihood returns a constant image there, which is what the engine takes
                none = likelihood.none_like(xT[:1])
                none_value = float(none.flatten()[0])
                if not bool((none == none_value).all()):
                    raise NotImplementedError("Amortized corrector steps need a constant likelihood.none_like()")
            return eps_model.network.engine().sample_ddpm(xT, ddpm.tables(), mode="amortized", condition=condition,
                                                          pad_value=none_value,
                                                          noise=_noise_tensor(noise, ddpm.Ns, xT), seed=seed,
                                                          use_graph=use_graph, n_corrector=n_corr,
                                                          corrector_delta=float(conditioning.delta))
        return sample

    if isinstance(conditioning, Replacement):
        pad = getattr(likelihood, "pad_value", None)
        if pad is None:
            raise TypeError("Replacement conditioning needs a likelihood with a pad_value (In/OutPainting)")

        @torch.no_grad()
        def sample(xT, condition):
            return eps_model.network.engine().sample_ddpm(
                xT, ddpm.tables(), mode="replacement", condition=condition, pad_value=float(pad),
                replace_below_step=int(ddpm.Ns * conditioning.start_fraction), noise_condition=bool(conditioning.noise),
                noise=_noise_tensor(noise, ddpm.Ns, xT), seed=seed, use_graph=use_graph,
                n_corrector=int(getattr(conditioning, "n_corrector", 0)), corrector_delta=float(conditioning.delta))
        return sample

    raise NotImplementedError(f"no engine sampler for conditioning {type(conditioning).__name__}")
