"""DDPM reverse-chain sampling over the engine.

Mirrors the reference's operator surface for this path:

* ``DDPM(Ns)`` buffers and algebra ........ AD/image_diffusion/sde_diffusion.py:107-244
* ``Amortized`` / ``Replacement`` ......... AD/image_diffusion/conditioning.py:12-63
* ``InPainting`` / ``OutPainting`` / ``HyperResolution`` ... AD/image_diffusion/likelihoods.py:39-146
* ``get_prior_sample_fn`` / ``get_conditional_sample_fn(eps_model, ddpm, conditioning, likelihood)``
  ........................................ AD/image_diffusion/sampling.py:50-75, 80-133, 209-260
* The eps network is ANY callable ``eps_model(xi, i)`` (type ``Network``, sde_diffusion.py:11), as in the
  reference: ``lambda xi, i: ema_network(xi, 1.0 * i / ddpm.Ns)`` (AD/experiments/main.py:140) works
  unchanged.  When the callable turns out to be an engine-backed U-Net evaluated at ``t = i / Ns`` (an
  ``EpsModel``, or a closure over one that a two-point probe confirms), the whole chain runs as one
  native call (``cfm_sample_ddpm``); otherwise the chain runs step by step with the native step
  kernels (``cfm_ddpm_step``) around the Python callable.  ``ReconstructionGuidance`` needs the U-Net's
  backward pass and is out of scope for this inference engine (raises NotImplementedError).
* Noise: the reference draws fresh ``torch.randn_like`` noise on every call.  Here every call of a
  returned ``sample()`` draws a new Philox seed from torch's global CPU generator (mixed with the
  rank under ``torch.distributed``), so successive calls and different ranks get independent noise;
  ``seed=`` / ``noise=`` pin it for reproducibility.

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


def int_b(t):
    return bm * t + (bd - bm) * t ** 2 / 2


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

    def model_time(self, device=None) -> torch.Tensor:
        """t fed to the U-Net at step i: ``1.0 * i / Ns`` on an int64 tensor (main.py:140), evaluated on ``device`` - where
        the reference's lambda evaluates it.  (torch's CUDA division by a Python scalar multiplies by the fp32 reciprocal,
        its CPU division divides: the two differ in the last bit for some i, and a fused chain must feed the U-Net exactly
        the t the caller's own ``eps_model(xi, i)`` would.)"""
        return ((1.0 * torch.arange(self.Ns, dtype=torch.long, device=device)) / self.Ns).cpu()

    def tables(self, device=None) -> dict:
        names = ("sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
                 "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2",
                 "posterior_log_variance_clipped")
        tb = {n: getattr(self, n).detach().cpu() for n in names}
        tb["model_time"] = self.model_time(device)
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

    def em_step(self, xi: torch.Tensor, noise_hat: torch.Tensor, i: int, z: Optional[torch.Tensor] = None, seed: int = 0):
        """One reverse-time Euler-Maruyama step of the VP SDE from a noise prediction (sampling.py:100-111 ``em_step`` with
        backward_drift / backward_diffusion / score_from_noise, sde_diffusion.py:170-205), as one native launch:
        ``x = xi - dt*drift + g*z*sqrt(dt)``, ``drift = -0.5*xi*xi - g^2*score`` (sic: the reference's ``drift`` multiplies by x,
        not beta_t - see oracle/ddpm.py), ``score = -noise_hat/sigma_t``,
        ``t = ts[i]``, ``dt = 1/Ns``.  ``z``: injected normals (else Philox(seed), stream i).  Returns a new tensor."""
        t = self.ts[int(i)]
        beta_t = beta(t)
        sigma_t = torch.sqrt(1 - torch.exp(-int_b(t)))
        x = xi.detach().to(torch.float32).contiguous().clone()
        eps = noise_hat.detach().to(torch.float32).contiguous()
        if x.device.type != "cuda" or eps.device != x.device or eps.shape != x.shape:
            raise ValueError("xi and noise_hat must be CUDA tensors of the same shape")
        zd = None if z is None else z.to(device=x.device, dtype=torch.float32).contiguous()
        lib = _lib.load()
        with torch.cuda.device(x.device):
            rc = lib.cfm_ddpm_em_step(C.c_void_p(x.data_ptr()), C.c_void_p(eps.data_ptr()), float(beta_t), float(sigma_t),
                                      1 / self.Ns, None if zd is None else C.c_void_p(zd.data_ptr()), C.c_uint64(seed),
                                      int(i), x.numel(), C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
        _lib.check(rc)
        return x

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
        """All (h, w) pairs in ONE call.  torch's CPU generator fills a tensor element by element in memory order, so
        ``randint(lo, hi, (B, 2))`` consumes the global stream exactly like the reference's per-sample loop of two scalar
        draws, h before w (likelihoods.py:49-53, 78-87; tests/test_host_logic.py pins this against the loop)."""
        lo, hi = 5, image_size - self.patch_size - 5
        if batch == 0:
            return torch.empty((0, 2), dtype=torch.int32)
        return torch.randint(lo, hi, size=(batch, 2)).to(torch.int32)

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


def resize_bilinear(images: torch.Tensor, size) -> torch.Tensor:
    """``F.interpolate(images, size, mode="bilinear", align_corners=False)`` as one native kernel (cfm_resize_bilinear)."""
    if images.device.type != "cuda":
        raise RuntimeError("images must be on a CUDA device")
    if images.dim() != 4:
        raise ValueError("images must be [B, C, H, W]")
    h, w = (int(size), int(size)) if isinstance(size, int) else (int(size[0]), int(size[1]))
    src = images.detach().to(torch.float32).contiguous()
    B, Cc, H, W = src.shape
    out = torch.empty((B, Cc, h, w), device=src.device, dtype=torch.float32)
    lib = _lib.load()
    with torch.cuda.device(src.device):
        rc = lib.cfm_resize_bilinear(C.c_void_p(out.data_ptr()), C.c_void_p(src.data_ptr()), B * Cc, H, W, h, w,
                                     C.c_void_p(torch.cuda.current_stream(src.device).cuda_stream))
    _lib.check(rc)
    return out


class HyperResolution(Likelihood):
    @classmethod
    def from_configdict(cls, config):
        return cls(config["target_height"], config["target_width"])

    def __init__(self, target_height: int, target_width: int):
        self.target_height, self.target_width = target_height, target_width

    def sample(self, images: torch.Tensor) -> torch.Tensor:
        lo = resize_bilinear(images, (self.target_height, self.target_width))
        return resize_bilinear(lo, (images.shape[2], images.shape[3]))

    def none_like(self, x):
        return torch.zeros_like(x)


def get_likelihood(type_: str):
    return {"inpainting": InPainting, "outpainting": OutPainting, "hyperresolution": HyperResolution}[type_.lower()]


def downsample_images(images, target_size):
    """mnist/utils_mnist_hy.py:18-28."""
    return resize_bilinear(images, target_size)


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


def _fresh_seed(seed: Optional[int]) -> int:
    """Explicit seed: reproducible.  None: a new seed per call from torch's global CPU generator (advances it, as the
    reference's randn_like calls advance the global generator), mixed with the rank under torch.distributed."""
    if seed is not None:
        return int(seed)
    s_ = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64))
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            s_ ^= (dist.get_rank() + 1) * 0x9E3779B97F4A7C15 & (2 ** 63 - 1)
    except Exception:
        pass
    return s_


def _find_unet(obj, depth: int = 0):
    """An engine-backed U-Net reachable from a callable: itself, its ``ema_model`` / ``module`` / ``network`` attribute
    (EMA and DataParallel wrappers), or a cell of its closure."""
    if isinstance(obj, UNetModel):
        return obj
    if depth > 2:
        return None
    for name in ("network", "ema_model", "module", "model"):
        sub = getattr(obj, name, None)
        if sub is not None and sub is not obj:
            r = _find_unet(sub, depth + 1)
            if r is not None:
                return r
    for cell in getattr(obj, "__closure__", None) or ():
        try:
            r = _find_unet(cell.cell_contents, depth + 1)
        except ValueError:
            r = None
        if r is not None:
            return r
    return None


def _engine_behind(eps_model, ddpm: DDPM, in_channels_needed: Optional[int] = None):
    """The U-Net whose engine can run the whole chain natively, or None.  A plain callable qualifies only if it is
    observably ``xi, i -> unet(xi, i / Ns)``: it is probed at two chain indices on a random input and must reproduce
    the engine's own forward bit for bit (the engine is deterministic)."""
    if isinstance(eps_model, EpsModel):
        return eps_model.network if isinstance(eps_model.network, UNetModel) else None
    net = _find_unet(eps_model)
    if net is None or net.num_classes is not None:
        return None
    try:
        p = next(net.parameters())
        if p.device.type != "cuda":
            return None
        cfg = net.config
        g = torch.Generator(device="cpu").manual_seed(12345)
        x = torch.randn(2, cfg.in_channels, cfg.image_size, cfg.image_size, generator=g).to(p.device)
        for idx in (0, ddpm.Ns - 1):
            i = torch.full((2,), idx, device=p.device, dtype=torch.long)
            got = eps_model(x, i)
            want = net.engine().forward(x, 1.0 * i / ddpm.Ns)
            if got.shape != want.shape or not torch.equal(got.to(torch.float32), want):
                return None
        return net
    except Exception:
        return None


def _check_condition(xT: torch.Tensor, condition: torch.Tensor) -> torch.Tensor:
    """The reference's ``torch.where`` / ``concat`` broadcast a [1, C, H, W] condition; the native chain indexes it
    element by element, so it is expanded here and anything else is rejected."""
    if condition.dim() != xT.dim() or tuple(condition.shape[1:]) != tuple(xT.shape[1:]) or condition.shape[0] not in (1, xT.shape[0]):
        raise ValueError(f"condition must be {tuple(xT.shape)} (or batch 1), got {tuple(condition.shape)}")
    if condition.device != xT.device:
        raise ValueError("condition and xT must live on the same device")
    return condition.expand_as(xT) if condition.shape[0] != xT.shape[0] else condition


def _tables_c(ddpm: DDPM):
    keep = []
    tb = _lib.DdpmTablesC()
    tb.Ns = ddpm.Ns
    for name, t in ddpm.tables().items():
        h = t.detach().to("cpu", torch.float32).contiguous()
        keep.append(h)
        setattr(tb, name, C.cast(h.data_ptr(), C.POINTER(C.c_float)))
    return tb, keep


def _stepwise_chain(eps_model, ddpm: DDPM, xT, mode: str, condition=None, pad_value=-2.0, replace_below_step=None,
                    noise_condition=True, noise=None, seed=0, n_corrector=0, corrector_delta=0.1, none_value=0.0):
    """The reverse chains of sampling.py:50-75 / 80-133 / 209-260 around an arbitrary Python eps network: every
    elementwise step is one native launch (cfm_ddpm_step), the network call stays with the caller's callable."""
    lib = _lib.load()
    if xT.device.type != "cuda":
        raise RuntimeError("xT must be on a CUDA device (no CPU fallback)")
    x = xT.detach().to(torch.float32).contiguous().clone()
    B = x.shape[0]
    Ns = ddpm.Ns
    cd = None if condition is None else condition.detach().to(torch.float32).contiguous()
    tb, keep = _tables_c(ddpm)
    opt = _lib.DdpmOptionsC()
    opt.mode = {"prior": _lib.DDPM_PRIOR, "replacement": _lib.DDPM_REPLACEMENT, "amortized": _lib.DDPM_AMORTIZED}[mode]
    opt.pad_value = float(pad_value)
    opt.replace_below_step = Ns if replace_below_step is None else int(replace_below_step)
    opt.noise_condition = int(noise_condition)
    opt.n_corrector = int(n_corrector)
    opt.corrector_delta = float(corrector_delta)
    nd = None
    if noise is not None:
        nd = noise.to(device=x.device, dtype=torch.float32).contiguous()
        assert nd.numel() == Ns * (2 + int(n_corrector)) * x.numel(), "noise must be [Ns, 2 + n_corrector, B*C*H*W]"
    if B == 0:
        return x
    none_cond = torch.full_like(x, none_value) if (mode == "amortized" and n_corrector) else None
    stream = lambda: C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())

    def launch(i, phase, eps):
        with torch.cuda.device(x.device):
            _lib.check(lib.cfm_ddpm_step(ptr(x), ptr(eps), ptr(cd) if mode == "replacement" else None, C.byref(tb), C.byref(opt),
                                         i, phase, ptr(nd), C.c_uint64(seed), x.numel(), stream()))

    def net(i, cond_for_net):
        idx = torch.full((B,), i, device=x.device, dtype=torch.long)
        inp = x if cond_for_net is None else torch.cat((x, cond_for_net), dim=-3)
        return eps_model(inp, idx).to(torch.float32).contiguous()

    for i in reversed(range(Ns)):
        if mode == "replacement":
            launch(i, 0, None)
        launch(i, 1, net(i, cd if mode == "amortized" else None))
        for c in range(int(n_corrector)):
            launch(i, 2 + c, net(i, none_cond if mode == "amortized" else None))
    del keep
    return x


def get_prior_sample_fn(eps_model, ddpm: DDPM, conditioning=None, likelihood=None, *, noise=None, seed: Optional[int] = None,
                        use_graph: bool = False) -> Callable:
    if not callable(eps_model):
        raise TypeError("eps_model must be callable as eps_model(xi, i)")
    amortized = isinstance(conditioning, Amortized)
    state = {}

    @torch.no_grad()
    def sample(xT):
        if "net" not in state:
            state["net"] = _engine_behind(eps_model, ddpm)
        net, s_ = state["net"], _fresh_seed(seed)
        cond = likelihood.none_like(xT) if amortized else None      # x0_model's stand-in condition (sampling.py:36-37)
        if net is None:
            return _stepwise_chain(eps_model, ddpm, xT, "amortized" if amortized else "prior", condition=cond,
                                   noise=_noise_tensor(noise, ddpm.Ns, xT), seed=s_)
        return net.engine().sample_ddpm(xT, ddpm.tables(xT.device), mode="amortized" if amortized else "prior", condition=cond,
                                        noise=_noise_tensor(noise, ddpm.Ns, xT), seed=s_, use_graph=use_graph)

    return sample


def get_conditional_sample_fn(eps_model, ddpm: DDPM, conditioning, likelihood, *, noise=None, seed: Optional[int] = None,
                              use_graph: bool = False) -> Callable:
    if not callable(eps_model):
        raise TypeError("eps_model must be callable as eps_model(xi, i)")
    if getattr(conditioning, "n_corrector", 0) and not isinstance(conditioning, (Replacement, Amortized)):
        raise NotImplementedError("Langevin corrector steps (n_corrector > 0) run on the engine for Replacement and "
                                  "Amortized conditioning only")
    state = {}

    def backend():
        if "net" not in state:
            state["net"] = _engine_behind(eps_model, ddpm)
        return state["net"]

    if isinstance(conditioning, Amortized):
        n_corr = int(getattr(conditioning, "n_corrector", 0))

        @torch.no_grad()
        def sample(xT, condition):
            condition = _check_condition(xT, condition)
            none_value = 0.0
            if n_corr:
                # the corrector's x0_model call has no condition -> likelihood.none_like(xi) (sampling.py:36-37, 116);
                # every reference likelihood returns a constant image there, which is what the engine takes
                none = likelihood.none_like(xT[:1])
                none_value = float(none.flatten()[0])
                if not bool((none == none_value).all()):
                    raise NotImplementedError("Amortized corrector steps need a constant likelihood.none_like()")
            kw = dict(mode="amortized", condition=condition, pad_value=none_value, noise=_noise_tensor(noise, ddpm.Ns, xT),
                      seed=_fresh_seed(seed), n_corrector=n_corr, corrector_delta=float(conditioning.delta))
            net = backend()
            if net is None:
                return _stepwise_chain(eps_model, ddpm, xT, none_value=none_value, **kw)
            return net.engine().sample_ddpm(xT, ddpm.tables(xT.device), use_graph=use_graph, **kw)
        return sample

    if isinstance(conditioning, Replacement):
        pad = getattr(likelihood, "pad_value", None)
        if pad is None:
            raise TypeError("Replacement conditioning needs a likelihood with a pad_value (In/OutPainting)")

        @torch.no_grad()
        def sample(xT, condition):
            condition = _check_condition(xT, condition)
            kw = dict(mode="replacement", condition=condition, pad_value=float(pad),
                      replace_below_step=int(ddpm.Ns * conditioning.start_fraction), noise_condition=bool(conditioning.noise),
                      noise=_noise_tensor(noise, ddpm.Ns, xT), seed=_fresh_seed(seed),
                      n_corrector=int(getattr(conditioning, "n_corrector", 0)), corrector_delta=float(conditioning.delta))
            net = backend()
            if net is None:
                return _stepwise_chain(eps_model, ddpm, xT, **kw)
            return net.engine().sample_ddpm(xT, ddpm.tables(xT.device), use_graph=use_graph, **kw)
        return sample

    raise NotImplementedError(f"no engine sampler for conditioning {type(conditioning).__name__}")
