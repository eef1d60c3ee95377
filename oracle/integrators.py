"""CPU oracle: ODE integrators used by the CFM sampling loops.  TEST INFRASTRUCTURE ONLY.

Neither integrator lives under /root/reference: they come from the un-vendored,
un-pinned dependencies ``torchdyn`` (fixed-step Euler behind
``NeuralODE.trajectory``; call sites cifar10/compute_fid.py:70,79,
cifar10/utils_cifar.py:34-39, mnist/utils_mnist2.py:125-134) and ``torchdiffeq``
(``odeint(..., method="dopri5")``; call sites cifar10/compute_fid.py:83-85,
mnist/utils_mnist.py:101-108, conditional_mnist.ipynb cell 4).  Their published
algorithms are restated here.  PARITY UNPINNED at this boundary: the reference
holds no golden vectors for either; the vectors in tests/golden/ are generated
from this restatement, and the analytic-ODE tests in tests/test_integrators.py
check it against closed-form solutions.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple, Union

import torch

State = Union[torch.Tensor, Tuple[torch.Tensor, ...]]


def euler_trajectory(f: Callable, x: torch.Tensor, t_span: torch.Tensor) -> torch.Tensor:
    """torchdyn fixed-step driver with the Euler tableau.

    ``t`` is advanced by accumulation (t <- t + dt) and the next ``dt`` is
    re-derived from ``t_span`` so the grid is re-synchronised every step; all in
    the dtype of ``t_span`` (fp32 at every call site).  Returns the stacked
    trajectory ``[len(t_span), *x.shape]`` including the initial state.
    """
    t = t_span[0]
    dt = t_span[1] - t_span[0]
    sol = [x]
    n = len(t_span) - 1
    for step in range(1, n + 1):
        x = x + dt * f(t, x)
        t = t + dt
        if step < n:
            dt = t_span[step + 1] - t
        sol.append(x)
    return torch.stack(sol)


def euler_cfg_trajectory(f_cond: Callable, f_uncond: Callable, w: float, x: torch.Tensor, t_span: torch.Tensor) -> torch.Tensor:
    """Classifier-free guidance on the Euler driver above (EXTENSION, not in the reference: SURVEY F7).

    Two evaluations per step, combined as ``v_c + w (v_c - v_u)`` in fp32, in that order of operations.
    """
    wt = torch.tensor(w, dtype=torch.float32)

    def f(t, xx):
        vc, vu = f_cond(t, xx), f_uncond(t, xx)
        return vc + wt * (vc - vu)

    return euler_trajectory(f, x, t_span)


def sde_euler_maruyama(drift: Callable, score, x: torch.Tensor, ts: torch.Tensor, dt: float, sigma: float, noise: Callable) -> torch.Tensor:
    """``torchsde.sdeint(sde, y0, ts, dt=dt)`` with the default scheme of a diagonal Ito SDE ("euler": Euler-Maruyama) for
    the SF2M sampler of conditional_mnist.ipynb cells 11-12: ``f = flow(t, y) + score(t, y)``, ``g = sigma``.
    torchsde is not under /root/reference and is not pinned: PARITY UNPINNED; the published scheme is restated -
    fixed steps ``dt`` from ts[0] to ts[-1] (last one shortened), ``y <- y + f dt + g dW`` with ``dW = sqrt(dt) z`` and
    z drawn by ``noise(shape)`` once per step (torchsde's BrownianInterval draws are replaced by injected normals)."""
    import math
    t0, t1 = float(ts[0]), float(ts[-1])
    n_steps = max(int(math.ceil((t1 - t0) / dt - 1e-9)), 0)
    for k in range(n_steps):
        t = t0 + k * dt
        h = torch.tensor(min(dt, t1 - t), dtype=torch.float32)
        tt = torch.tensor(t, dtype=torch.float32)
        f = drift(tt, x) if score is None else drift(tt, x) + score(tt, x)
        dw = noise(x.shape) * torch.sqrt(h)
        x = (x + f * h) + sigma * dw
    return x


def euler_time_grid(t_span: torch.Tensor) -> Tuple[List[float], List[float]]:
    """The (t_k, dt_k) pairs the loop above feeds to f / uses in the update."""
    t = t_span[0]
    dt = t_span[1] - t_span[0]
    ts, dts = [], []
    n = len(t_span) - 1
    for step in range(1, n + 1):
        ts.append(float(t)); dts.append(float(dt))
        t = t + dt
        if step < n:
            dt = t_span[step + 1] - t
    return ts, dts


# --- Dormand-Prince 5(4) -----------------------------------------------------------------

_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
_C_SOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
_C_ERR = [35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50043, 125 / 192 - 451 / 720,
          -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1.0 / 60.0]
_C_MID = [6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
          187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2]

DOPRI5_TABLEAU = dict(alpha=_ALPHA, beta=_BETA, c_sol=_C_SOL, c_err=_C_ERR, c_mid=_C_MID)


def _flatten(state: Sequence[torch.Tensor]) -> torch.Tensor:
    return torch.cat([s.reshape(-1) for s in state])


def _rms(v: torch.Tensor) -> float:
    return float(v.abs().pow(2).mean().sqrt())


def _mixed_norm(v: torch.Tensor, sizes: Sequence[int]) -> float:
    """max over tuple components of each component's RMS (torchdiffeq tuple-state rule)."""
    if len(sizes) == 1:
        return _rms(v)
    out, off = 0.0, 0
    for n in sizes:
        out = max(out, _rms(v[off:off + n])); off += n
    return out


def dopri5(f: Callable, y0: State, t: Sequence[float], rtol: float, atol: float,
           max_steps: int = 100000, stats: dict | None = None) -> State:
    """Adaptive RK5(4) with FSAL, dense output at the requested times.

    Controller constants: safety 0.9, ifactor 10, dfactor 0.2, order 5; initial
    step from the Hairer heuristic with order 4; accept when the (mixed) RMS norm of
    err / (atol + rtol*max(|y0|,|y1|)) <= 1; steps are NOT clipped to the output
    times - outputs come from the 4th-order interpolant.  Time arithmetic is
    float64 on the host; state arithmetic stays in the state's dtype.
    Returns the solution stacked over ``t`` (tuple in -> tuple out).
    """
    is_tuple = isinstance(y0, (tuple, list))
    comps = list(y0) if is_tuple else [y0]
    shapes = [c.shape for c in comps]
    sizes = [c.numel() for c in comps]

    def unflat(v):
        out, off = [], 0
        for s, n in zip(shapes, sizes):
            out.append(v[off:off + n].reshape(s)); off += n
        return tuple(out)

    nfe = 0

    def func(tt: float, v: torch.Tensor) -> torch.Tensor:
        nonlocal nfe
        nfe += 1
        tt_t = torch.tensor(tt, dtype=comps[0].dtype)
        if is_tuple:
            return _flatten(f(tt_t, unflat(v)))
        return f(tt_t, v.reshape(shapes[0])).reshape(-1)

    norm = lambda v: _mixed_norm(v, sizes)
    y = _flatten(comps)
    t0 = float(t[0])
    f0 = func(t0, y)
    # initial step (order = 4)
    scale = atol + y.abs() * rtol
    d0, d1 = norm(y / scale), norm(f0 / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    f1 = func(t0 + h0, y + h0 * f0)
    d2 = norm((f1 - f0) / scale) / h0
    h1 = max(1e-6, h0 * 1e-3) if (d1 <= 1e-15 and d2 <= 1e-15) else (0.01 / max(d1, d2)) ** (1.0 / 5.0)
    dt = min(100 * h0, h1)

    t_lo, t_hi = t0, t0
    y_lo = y
    coeffs = None
    outputs = [y]
    n_steps = n_accept = 0
    for t_out in [float(v) for v in t[1:]]:
        while t_out > t_hi:
            assert n_steps < max_steps, "max_num_steps exceeded"
            n_steps += 1
            k = [f0]
            for a, brow in zip(_ALPHA, _BETA):
                yi = y + dt * sum(b * kj for b, kj in zip(brow, k) if b != 0)
                k.append(func(t_hi + a * dt, yi))
            y1 = yi                      # FSAL: last stage is the 5th-order solution
            err = dt * sum(c * kj for c, kj in zip(_C_ERR, k) if c != 0)
            tol = atol + rtol * torch.max(y.abs(), y1.abs())
            ratio = norm(err / tol)
            if ratio <= 1:
                n_accept += 1
                y_mid = y + dt * sum(c * kj for c, kj in zip(_C_MID, k) if c != 0)
                fa, fb = k[0], k[-1]
                coeffs = (2 * dt * (fb - fa) - 8 * (y1 + y) + 16 * y_mid,
                          dt * (5 * fa - 3 * fb) + 18 * y + 14 * y1 - 32 * y_mid,
                          dt * (fb - 4 * fa) - 11 * y - 5 * y1 + 16 * y_mid,
                          dt * fa, y)
                t_lo, t_hi = t_hi, t_hi + dt
                y_lo, y, f0 = y, y1, k[-1]
            # step-size update (order 5)
            if ratio == 0:
                factor = 10.0
            else:
                dfac = 1.0 if ratio < 1 else 0.2
                factor = min(10.0, max(0.9 / ratio ** 0.2, dfac))
            dt = dt * factor
        x = (t_out - t_lo) / (t_hi - t_lo)
        a_, b_, c_, d_, e_ = coeffs
        outputs.append((((a_ * x + b_) * x + c_) * x + d_) * x + e_)
    if stats is not None:
        stats.update(nfe=nfe, steps=n_steps, accepted=n_accept)
    stacked = torch.stack(outputs)
    if is_tuple:
        out, off = [], 0
        for s, n in zip(shapes, sizes):
            out.append(stacked[:, off:off + n].reshape(len(outputs), *s)); off += n
        return tuple(out)
    return stacked.reshape(len(outputs), *shapes[0])
