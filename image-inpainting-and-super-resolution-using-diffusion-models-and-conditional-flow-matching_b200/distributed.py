"""Multi-GPU sharding of the sampling loop: one process per GPU, no collective inside the loop.

Samples are independent (GroupNorm and attention are per-sample), so the batch is split across
ranks; each rank runs the fused loop on its own engine with replicated weights.  NCCL (or gloo
in CPU tests) is used only after the loop, to gather the finished uint8 images.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of `total` samples owned by `rank`; sizes differ by at most 1."""
    if world <= 0 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_uint8(local: torch.Tensor, total: int, group=None) -> Optional[torch.Tensor]:
    """All-gather ragged uint8 shards [n_r, C, H, W] into [total, C, H, W] (every rank gets the result)."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    max_n = max(hi - lo for lo, hi in sizes)
    if total % world == 0:      # equal shards: one collective straight into the result, no padding or concatenation
        out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = torch.zeros((max_n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)


def sample_euler_sharded(model, x0_global: torch.Tensor, t_span: torch.Tensor, y_global=None, cond_global=None,
                         use_graph: bool = True, group=None):
    """Each rank integrates its slice of `x0_global`; returns (local final state, gathered uint8 images)."""
    from .integrators import sample_euler
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(x0_global.shape[0], rank, world)
    dev = next(model.parameters()).device
    x0 = x0_global[lo:hi].to(dev)
    y = None if y_global is None else y_global[lo:hi].to(dev)
    cond = None if cond_global is None else cond_global[lo:hi].to(dev)
    x, img = sample_euler(model, x0, t_span, y=y, cond=cond, return_uint8=True, use_graph=use_graph)
    return x, gather_uint8(img, x0_global.shape[0], group)


def odeint_sharded(func, y0_global: torch.Tensor, t: torch.Tensor, rtol: float = 1e-5, atol: float = 1e-5, group=None,
                   stats: Optional[dict] = None):
    """dopri5 over a batch-sharded state with ONE step controller: every rank integrates its slice of `y0_global`
    and the error norms are all-reduced (a few doubles per attempted step), so all ranks accept / reject the same
    steps a single process would on the whole batch (torchdiffeq's norm is global over the state).
    Returns the local trajectory [len(t), n_local, ...]."""
    from .integrators import odeint
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(y0_global.shape[0], rank, world)
    g = (group if group is not None else dist.group.WORLD) if dist.is_initialized() else None
    return odeint(func, y0_global[lo:hi].cuda(), t, rtol=rtol, atol=atol, method="dopri5", stats=stats, norm_group=g)
