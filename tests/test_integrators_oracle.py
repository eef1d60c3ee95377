"""Oracle integrators against closed-form ODE solutions (CPU)."""
import math

import torch

from oracle import integrators as I


def test_euler_grid_and_trajectory():
    t_span = torch.linspace(0, 1, 101)
    ts, dts = I.euler_time_grid(t_span)
    assert len(ts) == 100 and abs(sum(dts) - 1.0) < 1e-6 and ts[0] == 0.0
    traj = I.euler_trajectory(lambda t, x: -x, torch.ones(3), t_span)
    assert traj.shape == (101, 3)
    assert abs(float(traj[-1, 0]) - (1 - 0.01) ** 100) < 1e-5
    # training-time variant: linspace(0,1,100) is 99 steps (utils_cifar.py:34-41)
    assert len(I.euler_time_grid(torch.linspace(0, 1, 100))[0]) == 99


def test_dopri5_exponential_and_tuple_state():
    st = {}
    y = I.dopri5(lambda t, x: -2.0 * x, torch.ones(4), [0.0, 1.0], 1e-6, 1e-6, stats=st)
    assert abs(float(y[-1, 0]) - math.exp(-2.0)) < 1e-5
    assert st["nfe"] == 2 + 6 * st["steps"]
    # tuple state with the reference's "conditioning drifts as e^t" behaviour
    ya, yb = I.dopri5(lambda t, s: (s[1], s[1]), (torch.zeros(2), torch.ones(2)), [0.0, 1.0], 1e-6, 1e-6)
    assert abs(float(yb[-1, 0]) - math.e) < 1e-4 and abs(float(ya[-1, 0]) - (math.e - 1)) < 1e-4


def test_dopri5_time_dependent():
    y = I.dopri5(lambda t, x: torch.cos(t) * torch.ones_like(x), torch.zeros(2), [0.0, 0.5, 2.0], 1e-7, 1e-7)
    assert abs(float(y[1, 0]) - math.sin(0.5)) < 1e-5 and abs(float(y[2, 0]) - math.sin(2.0)) < 1e-5
