"""ODE integrators over the engine: the seams ``NeuralODE(model, solver).trajectory(x, t_span)``
(torchdyn; cifar10/compute_fid.py:70,79) and ``odeint(f, y0, t, rtol=, atol=, method=)``
(torchdiffeq; cifar10/compute_fid.py:83-85, mnist/utils_mnist.py:101-108).

* Euler with an engine-backed model runs as ONE native call (``cfm_sample_euler``): the whole
  fixed-step loop - U-Net, update, optional trajectory store and uint8 conversion - stays on
  the device, optionally as a single CUDA graph.
* dopri5 keeps the accept/reject controller on the host (one device->host read of the error
  norm per attempted step, as the reference does) and runs the stage combinations and the
  error norm in fused native kernels (``cfm_rk_combine`` / ``cfm_rk_error_sumsq``).
* Any other callable ``f(t, x)`` is integrated with the same native state kernels.
"""
from __future__ import annotations

import ctypes as C
import inspect
import math
from typing import Callable, List, Optional, Sequence, Tuple, Union

import torch

from . import _lib
from .models import UNetModel, UNetModelWrapper, InPaintModelWrapper, SuperResModelWrapper

_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
_C_ERR = [35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50043, 125 / 192 - 451 / 720,
          -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1.0 / 60.0]
_C_MID = [6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
          187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2]


def euler_time_grid(t_span: torch.Tensor) -> Tuple[List[float], List[float]]:
    """(t_k, dt_k) of torchdyn's fixed-step driver: t accumulates, dt is re-derived from t_span, in t_span's dtype."""
    ts_ = t_span.detach().to("cpu")
    t = ts_[0]
    dt = ts_[1] - ts_[0]
    ts, dts = [], []
    n = len(ts_) - 1
    for step in range(1, n + 1):
        ts.append(float(t)); dts.append(float(dt))
        t = t + dt
        if step < n:
            dt = ts_[step + 1] - t
    return ts, dts


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check_state(*tensors):
    """The native state kernels take raw pointers: fp32, contiguous, one CUDA device, one element count."""
    ref = tensors[0]
    for t in tensors:
        if not torch.is_tensor(t) or t.device.type != "cuda" or t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("state tensors must be contiguous fp32 CUDA tensors")
        if t.device != ref.device or t.numel() != ref.numel():
            raise ValueError("state tensors must share device and element count")


def rk_combine(y: torch.Tensor, ks: Sequence[torch.Tensor], coefs: Sequence[float], dt: float,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = y + dt * sum_j coefs[j] * ks[j]   (one fused kernel; zero coefficients are dropped)."""
    lib = _lib.load()
    pairs = [(k, c) for k, c in zip(ks, coefs) if c != 0]
    assert 1 <= len(pairs) <= 8
    _check_state(y, *[k for k, _ in pairs], *([] if out is None else [out]))
    if out is None:
        out = torch.empty_like(y)
    kp = (C.c_void_p * len(pairs))(*[k.data_ptr() for k, _ in pairs])
    cf = (C.c_float * len(pairs))(*[float(c) for _, c in pairs])
    with torch.cuda.device(y.device):
        rc = lib.cfm_rk_combine(C.c_void_p(out.data_ptr()), C.c_void_p(y.data_ptr()), kp, cf, len(pairs), float(dt),
                                y.numel(), _stream(y.device))
    _lib.check(rc)
    return out


def rk_error_sumsq(y0, y1, ks, coefs, dt, rtol, atol, scratch: torch.Tensor) -> torch.Tensor:
    """Device scalar (float64) = sum((dt*sum_j c_j k_j / (atol + rtol*max(|y0|,|y1|)))^2)."""
    lib = _lib.load()
    pairs = [(k, c) for k, c in zip(ks, coefs) if c != 0]
    _check_state(y0, y1, *[k for k, _ in pairs])
    if scratch.device != y0.device or scratch.dtype != torch.float64 or scratch.numel() < 1:
        raise ValueError("scratch must be a float64 tensor on the state's device")
    kp = (C.c_void_p * len(pairs))(*[k.data_ptr() for k, _ in pairs])
    cf = (C.c_float * len(pairs))(*[float(c) for _, c in pairs])
    with torch.cuda.device(y0.device):
        rc = lib.cfm_rk_error_sumsq(C.c_void_p(scratch.data_ptr()), C.c_void_p(y0.data_ptr()), C.c_void_p(y1.data_ptr()),
                                    kp, cf, len(pairs), float(dt), float(rtol), float(atol), y0.numel(), _stream(y0.device))
    _lib.check(rc)
    return scratch


def rk_scaled_sumsq(a, b, y, rtol, atol, scratch: torch.Tensor) -> torch.Tensor:
    """Device scalar (float64) = sum(((a - b) / (atol + rtol*|y|))^2); ``b`` may be None (initial-step heuristic)."""
    lib = _lib.load()
    _check_state(a, y, *([] if b is None else [b]))
    if scratch.device != y.device or scratch.dtype != torch.float64 or scratch.numel() < 1:
        raise ValueError("scratch must be a float64 tensor on the state's device")
    with torch.cuda.device(y.device):
        rc = lib.cfm_rk_scaled_sumsq(C.c_void_p(scratch.data_ptr()), C.c_void_p(a.data_ptr()),
                                     None if b is None else C.c_void_p(b.data_ptr()), C.c_void_p(y.data_ptr()),
                                     float(rtol), float(atol), y.numel(), _stream(y.device))
    _lib.check(rc)
    return scratch


def rk_dense_output(y0, y1, y_mid, f0, f1, dt: float, x: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Quartic dense output of an accepted dopri5 step at fraction ``x`` of the step (one fused kernel)."""
    lib = _lib.load()
    _check_state(y0, y1, y_mid, f0, f1, *([] if out is None else [out]))
    if out is None:
        out = torch.empty_like(y0)
    with torch.cuda.device(y0.device):
        rc = lib.cfm_rk_dense_output(C.c_void_p(out.data_ptr()), C.c_void_p(y0.data_ptr()), C.c_void_p(y1.data_ptr()),
                                     C.c_void_p(y_mid.data_ptr()), C.c_void_p(f0.data_ptr()), C.c_void_p(f1.data_ptr()),
                                     float(dt), float(x), y0.numel(), _stream(y0.device))
    _lib.check(rc)
    return out


class NeuralODE(torch.nn.Module):
    """``NeuralODE(vector_field, solver="euler"|"dopri5", atol=, rtol=).trajectory(x, t_span)``."""

    def __init__(self, vector_field, solver: str = "euler", sensitivity: str = "adjoint", atol: float = 1e-4,
                 rtol: float = 1e-4, use_graph: bool = False, **_unused):
        super().__init__()
        if solver not in ("euler", "dopri5"):
            raise NotImplementedError(f"solver {solver!r}: the engine implements 'euler' and 'dopri5'")
        self.vf = vector_field
        self.solver, self.atol, self.rtol, self.use_graph = solver, atol, rtol, use_graph
        self._takes_args = False
        if not isinstance(vector_field, UNetModel):
            try:
                self._takes_args = "args" in inspect.signature(vector_field).parameters
            except (TypeError, ValueError):
                pass

    def _call(self, t, x):
        return self.vf(t, x, args={}) if self._takes_args else self.vf(t, x)

    @torch.no_grad()
    def trajectory(self, x: torch.Tensor, t_span: torch.Tensor) -> torch.Tensor:
        if self.solver == "dopri5":
            _warn_bf16_tolerance(self.vf, self.rtol, self.atol)
            return odeint(self._call, x, t_span, rtol=self.rtol, atol=self.atol, method="dopri5")
        ts, dts = euler_time_grid(t_span)
        if type(self.vf) is UNetModelWrapper and self.vf.num_classes is None:
            _, traj, _ = self.vf.engine().sample_euler(x, ts, dts, return_trajectory=True, use_graph=self.use_graph)
            return traj
        # generic callable: native update kernel around a Python-level vector field
        sol = [x]
        for t, dt in zip(ts, dts):
            v = self._call(torch.tensor(t, dtype=t_span.dtype, device=x.device), x)
            x = rk_combine(x.contiguous(), [v.contiguous()], [1.0], dt)
            sol.append(x)
        return torch.stack(sol)

    def forward(self, x, t_span):
        return t_span, self.trajectory(x, t_span)


def sample_euler(model: UNetModel, x0: torch.Tensor, t_span: torch.Tensor, y=None, cond=None, cond_drift=False,
                 return_uint8: bool = False, use_graph: bool = True, guidance_weight: Optional[float] = None):
    """Final state only (no 101x trajectory): what cifar10/compute_fid.py:73-88 actually consumes.

    Returns x_final, or (x_final, uint8 image) with ``return_uint8``.
    """
    ts, dts = euler_time_grid(t_span)
    x, _, img = model.engine().sample_euler(x0, ts, dts, y=y, cond=cond, cond_drift=cond_drift,
                                            return_uint8=return_uint8, use_graph=use_graph, guidance_weight=guidance_weight)
    return (x, img) if return_uint8 else x


@torch.no_grad()
def sample_sde(drift, score, x0: torch.Tensor, ts: torch.Tensor, dt: float, sigma: float = 0.1, y=None,
               noise: Optional[torch.Tensor] = None, seed: int = 0) -> torch.Tensor:
    """Euler-Maruyama sampling of ``dx = (drift(t,x,y) + score(t,x,y)) dt + sigma dW`` from ``ts[0]`` to ``ts[-1]`` with a
    fixed step ``dt`` - what ``torchsde.sdeint(SDE(model, score_model, labels), x0, ts=linspace(0, 1, 2), dt=0.01)`` does
    in conditional_mnist.ipynb cells 11-12 (scheme "euler", diagonal Ito noise, g = sigma).  Returns the final state.

    Two engine-backed U-Nets (``score`` may be None) run as ONE native call (cfm_sample_sde); any other callables are
    stepped with the native update kernel (cfm_sde_em_step).  ``noise``: [n_steps, *x0.shape] injected normals, else a
    Philox stream per step from ``seed``."""
    lib = _lib.load()
    t0, t1 = float(ts[0]), float(ts[-1])
    n_steps = max(int(math.ceil((t1 - t0) / dt - 1e-9)), 0)
    tg = [t0 + k * dt for k in range(n_steps)]
    dg = [min(dt, t1 - t) for t in tg]
    x = x0.detach().to(torch.float32).contiguous().clone()
    if x.device.type != "cuda":
        raise RuntimeError("x0 must be on a CUDA device (no CPU fallback)")
    B = x.shape[0]
    nd = None
    if noise is not None:
        nd = noise.to(device=x.device, dtype=torch.float32).contiguous()
        assert nd.numel() == n_steps * x.numel(), "noise must be [n_steps, *x0.shape]"
    if B == 0 or n_steps == 0:
        return x
    native = isinstance(drift, UNetModel) and (score is None or isinstance(score, UNetModel))
    if native:
        e_d = drift.engine()
        e_s = None if score is None else score.engine()
        yd = e_d._labels(y, B)
        tgc = (C.c_float * n_steps)(*tg)
        dgc = (C.c_float * n_steps)(*dg)
        with torch.cuda.device(x.device):
            rc = lib.cfm_sample_sde(e_d._h, None if e_s is None else e_s._h, B, C.c_void_p(x.data_ptr()),
                                    None if yd is None else C.c_void_p(yd.data_ptr()), tgc, dgc, n_steps, float(sigma),
                                    None if nd is None else C.c_void_p(nd.data_ptr()), C.c_uint64(seed), _stream(x.device))
        _lib.check(rc, e_d._h)
        return x
    call = (lambda f, t: f(t, x) if y is None else f(t, x, y))
    for k, (t, h) in enumerate(zip(tg, dg)):
        tt = torch.tensor(t, dtype=torch.float32, device=x.device)
        v = call(drift, tt).to(torch.float32).contiguous()
        s_ = None if score is None else call(score, tt).to(torch.float32).contiguous()
        zk = None if nd is None else nd.view(n_steps, -1)[k]
        with torch.cuda.device(x.device):
            rc = lib.cfm_sde_em_step(C.c_void_p(x.data_ptr()), C.c_void_p(v.data_ptr()), None if s_ is None else C.c_void_p(s_.data_ptr()),
                                     float(h), float(sigma), None if zk is None else C.c_void_p(zk.data_ptr()), C.c_uint64(seed), k,
                                     x.numel(), _stream(x.device))
        _lib.check(rc)
    return x


State = Union[torch.Tensor, Tuple[torch.Tensor, ...]]


def _warn_bf16_tolerance(func, rtol: float, atol: float) -> None:
    """The reference integrates an fp32 network at atol = rtol = 1e-4 (compute_fid.py:83-85).  A bf16 vector field carries
    ~4e-3 of relative rounding noise per evaluation: at tolerances tighter than that the step controller ends up
    resolving the noise (more rejected steps) - say so instead of silently spending NFEs."""
    model = getattr(func, "vf", func)
    if getattr(model, "precision", None) == "bf16" and min(rtol, atol) < 1e-3:
        import warnings
        warnings.warn(f"dopri5 with rtol={rtol:g}, atol={atol:g} on a precision='bf16' model: the vector field's rounding "
                      "noise (~4e-3 relative) exceeds the tolerance; build the model with precision='fp32' for adaptive "
                      "integration at this tolerance", RuntimeWarning, stacklevel=3)


def _allreduce_sum(values: Sequence[float], group) -> List[float]:
    """Sum a few host scalars over the ranks of ``group`` (NCCL: on the device; gloo: on the host)."""
    import torch.distributed as dist
    buf = torch.tensor(list(values), dtype=torch.float64)
    if dist.get_backend(group) == "nccl":
        buf = buf.cuda()
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return [float(v) for v in buf.cpu()]


@torch.no_grad()
def odeint(func: Callable, y0: State, t: torch.Tensor, rtol: float = 1e-7, atol: float = 1e-9,
           method: str = "dopri5", options=None, stats: Optional[dict] = None, norm_group=None) -> State:
    """``torchdiffeq.odeint`` for method in {"dopri5", "euler"}; tuple state in -> tuple out.

    ``norm_group`` (extension for batch-sharded sampling): a ``torch.distributed`` process group over which the
    error / initial-step norms are summed, so every rank takes the step sequence a single process would take on the
    whole batch (one small all-reduce per attempted step)."""
    if method == "euler":
        assert torch.is_tensor(y0)
        return NeuralODE(func, "euler").trajectory(y0, t)
    if method != "dopri5":
        raise NotImplementedError(method)
    _warn_bf16_tolerance(func, rtol, atol)
    is_tuple = isinstance(y0, (tuple, list))
    comps = [c.to(torch.float32).contiguous() for c in (y0 if is_tuple else [y0])]
    dev = comps[0].device
    if dev.type != "cuda":
        raise RuntimeError("odeint state must live on a CUDA device")
    shapes = [c.shape for c in comps]
    sizes = [c.numel() for c in comps]
    offs = [0]
    for n in sizes:
        offs.append(offs[-1] + n)

    def views(flat):
        return tuple(flat[offs[i]:offs[i + 1]].view(shapes[i]) for i in range(len(sizes)))

    nfe = 0

    def f(tt: float, flat: torch.Tensor) -> torch.Tensor:
        nonlocal nfe
        nfe += 1
        tt_t = torch.tensor(tt, dtype=torch.float32, device=dev)
        out = func(tt_t, views(flat) if is_tuple else flat.view(shapes[0]))
        if is_tuple:
            return torch.cat([o.reshape(-1).to(torch.float32) for o in out])
        return out.reshape(-1).to(torch.float32).contiguous()

    scratch = torch.zeros(1, dtype=torch.float64, device=dev)

    # element counts of each state component over all ranks that share the step controller
    gsizes = [float(n) for n in sizes]
    if norm_group is not None:
        gsizes = _allreduce_sum(gsizes, norm_group)

    def mixed_norm_of_ratio(y_a, y_b, ks, coefs, dt) -> float:
        """max over components of RMS(err/tol) - one fused kernel + one host read per component."""
        sums = []
        for i in range(len(sizes)):
            sl = slice(offs[i], offs[i + 1])
            s = rk_error_sumsq(y_a[sl], y_b[sl], [k[sl] for k in ks], coefs, dt, rtol, atol, scratch)
            sums.append(float(s.item()))
        if norm_group is not None:
            sums = _allreduce_sum(sums, norm_group)
        return max((s / n) ** 0.5 for s, n in zip(sums, gsizes))

    def scaled_norm(a: torch.Tensor, b: Optional[torch.Tensor], yy: torch.Tensor) -> float:
        """max over components of RMS((a - b) / (atol + rtol*|y|)) - the initial-step heuristic's norms, one fused
        deterministic kernel + one host read per component."""
        sums = []
        for i in range(len(sizes)):
            sl = slice(offs[i], offs[i + 1])
            s = rk_scaled_sumsq(a[sl], None if b is None else b[sl], yy[sl], rtol, atol, scratch)
            sums.append(float(s.item()))
        if norm_group is not None:
            sums = _allreduce_sum(sums, norm_group)
        return max((s / n) ** 0.5 for s, n in zip(sums, gsizes))

    y = torch.cat([c.reshape(-1) for c in comps])
    t_list = [float(v) for v in t]
    t0 = t_list[0]
    f0 = f(t0, y)
    d0, d1 = scaled_norm(y, None, y), scaled_norm(f0, None, y)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    f1 = f(t0 + h0, rk_combine(y, [f0], [1.0], h0))
    d2 = scaled_norm(f1, f0, y) / h0
    h1 = max(1e-6, h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** (1.0 / 5.0)
    dt = min(100 * h0, h1)

    t_lo = t_hi = t0
    coeffs = None
    outputs = [y]
    n_steps = n_accept = 0
    for t_out in t_list[1:]:
        while t_out > t_hi:
            n_steps += 1
            if n_steps > 100000:
                raise RuntimeError("max_num_steps exceeded")
            k = [f0]
            yi = None
            for a, brow in zip(_ALPHA, _BETA):
                yi = rk_combine(y, k, brow, dt)
                k.append(f(t_hi + a * dt, yi))
            y1 = yi
            ratio = mixed_norm_of_ratio(y, y1, k, _C_ERR, dt)
            if ratio <= 1:
                n_accept += 1
                # dense output of this step: (y0, y1, y_mid, f0, f1, dt) - evaluated by one kernel per output time
                coeffs = (y, y1, rk_combine(y, k, _C_MID, dt), k[0], k[-1], dt)
                t_lo, t_hi = t_hi, t_hi + dt
                y, f0 = y1, k[-1]
            if ratio == 0:
                factor = 10.0
            else:
                factor = min(10.0, max(0.9 / ratio ** 0.2, 1.0 if ratio < 1 else 0.2))
            dt = dt * factor
        xx = (t_out - t_lo) / (t_hi - t_lo)
        outputs.append(rk_dense_output(*coeffs, xx))
    if stats is not None:
        stats.update(nfe=nfe, steps=n_steps, accepted=n_accept)
    stacked = torch.stack(outputs)
    if is_tuple:
        return tuple(stacked[:, offs[i]:offs[i + 1]].reshape(len(outputs), *shapes[i]) for i in range(len(sizes)))
    return stacked.reshape(len(outputs), *shapes[0])
