"""FID sufficient statistics and the Frechet distance.

The reference scores samples with FID through ``cleanfid.fid.compute_fid(gen=gen_1_img, batch_size=1024, num_gen=50000)``
(cifar10/compute_fid.py:92-100) and ``torchmetrics.image.fid.FrechetInceptionDistance(feature=2048)``
(AD/experiments/main.py:261-267, 292-293).  Both are: features [N, D] of a fixed extractor -> mean and covariance of
the generated and the reference set -> ``|mu1 - mu2|^2 + tr(S1 + S2 - 2 (S1 S2)^(1/2))``.

What runs here:

* ``FIDStatistics.update(features)`` - running fp64 sums ``sum f`` and ``sum f f^T`` on the device, one native kernel
  (``cfm_fid_accumulate``), deterministic, batch by batch, so a 50 000-sample run never holds more than one batch.
* ``FIDStatistics.all_reduce()`` - the sums are additive over ranks: one ``all_reduce`` of D + D^2 doubles (33.6 MB at
  D = 2048) over NCCL after the sampling loop; the only other collective of the path besides the image gather.
* ``frechet_distance`` - on the device in fp64 (``tr (S1 S2)^(1/2)`` = sum of the square roots of the eigenvalues of
  ``S1 S2``, which are real and non-negative for PSD factors; cuSOLVER through ``torch.linalg`` - a plain library call).
* ``compute_fid(gen, feature_fn, num_gen, batch_size, reference)`` - the ``fid.compute_fid`` loop shape.

The feature extractor is a hook (``feature_fn(uint8 images [B,3,H,W]) -> [B, D]``): the Inception-v3 pool3 weights
that cleanfid / torchmetrics download are not available offline, and they are not part of the hot path.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Tuple

import torch

from . import _lib


class FIDStatistics:
    """Running ``n``, ``sum f`` and ``sum f f^T`` (fp64, on the device) of feature rows."""

    def __init__(self, dim: int, device=None):
        if not torch.cuda.is_available():
            raise _lib.EngineError("no CUDA device: the FID accumulation has no CPU fallback")
        self.dim = int(dim)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n = 0
        self.sum = torch.zeros(self.dim, dtype=torch.float64, device=self.device)
        self.outer = torch.zeros(self.dim, self.dim, dtype=torch.float64, device=self.device)

    def update(self, features: torch.Tensor) -> "FIDStatistics":
        if features.dim() != 2 or features.shape[1] != self.dim:
            raise ValueError(f"features must be [N, {self.dim}], got {tuple(features.shape)}")
        f = features.detach().to(device=self.device, dtype=torch.float32).contiguous()
        if f.shape[0] == 0:
            return self
        lib = _lib.load()
        with torch.cuda.device(self.device):
            rc = lib.cfm_fid_accumulate(C.c_void_p(self.sum.data_ptr()), C.c_void_p(self.outer.data_ptr()),
                                        C.c_void_p(f.data_ptr()), f.shape[0], self.dim,
                                        C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        _lib.check(rc)
        self.n += int(f.shape[0])
        return self

    def all_reduce(self, group=None) -> "FIDStatistics":
        """Sum the statistics over the ranks of ``group`` (NCCL on the device; a no-op without torch.distributed)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return self
        cnt = torch.tensor([float(self.n)], dtype=torch.float64, device=self.device)
        if dist.get_backend(group) != "nccl":       # gloo (CPU tests): reduce host copies
            bufs = [cnt.cpu(), self.sum.cpu(), self.outer.cpu()]
            for b in bufs:
                dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group)
            cnt, self.sum, self.outer = bufs[0].to(self.device), bufs[1].to(self.device), bufs[2].to(self.device)
        else:
            for b in (cnt, self.sum, self.outer):
                dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group)
        self.n = int(round(float(cnt)))
        return self

    def mean_cov(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """``mu = sum / n`` and the unbiased covariance ``(sum f f^T - n mu mu^T) / (n - 1)`` (``np.cov(rowvar=False)``)."""
        if self.n < 2:
            raise ValueError("need at least two feature rows")
        mu = self.sum / self.n
        cov = (self.outer - self.n * torch.outer(mu, mu)) / (self.n - 1)
        return mu, cov


def frechet_distance(mu1: torch.Tensor, sigma1: torch.Tensor, mu2: torch.Tensor, sigma2: torch.Tensor) -> float:
    """``|mu1 - mu2|^2 + tr(S1) + tr(S2) - 2 tr((S1 S2)^(1/2))`` in fp64 (cleanfid ``frechet_distance``)."""
    mu1, mu2 = mu1.to(torch.float64), mu2.to(torch.float64)
    s1, s2 = sigma1.to(torch.float64), sigma2.to(torch.float64)
    diff = mu1 - mu2
    ev = torch.linalg.eigvals(s1 @ s2)
    tr_sqrt = torch.sqrt(torch.clamp(ev.real, min=0.0)).sum()
    return float(diff.dot(diff) + torch.trace(s1) + torch.trace(s2) - 2.0 * tr_sqrt)


def compute_fid(gen: Callable, feature_fn: Callable, num_gen: int, batch_size: int,
                reference: Tuple[torch.Tensor, torch.Tensor], dim: Optional[int] = None, device=None, group=None) -> float:
    """The ``fid.compute_fid(gen=..., batch_size=..., num_gen=...)`` loop (cifar10/compute_fid.py:92-100): ``gen(z)``
    returns a batch of uint8 images, ``feature_fn`` maps them to [B, D] features, the statistics accumulate on the
    device, ranks are summed, and the distance to the ``reference`` (mu, sigma) is returned."""
    stats = None
    done = 0
    while done < num_gen:
        imgs = gen(None)
        take = min(int(imgs.shape[0]), num_gen - done)
        feats = feature_fn(imgs[:take])
        if stats is None:
            stats = FIDStatistics(dim or int(feats.shape[1]), device=device or feats.device)
        stats.update(feats)
        done += take
    stats.all_reduce(group)
    mu, cov = stats.mean_cov()
    return frechet_distance(mu, cov, reference[0].to(mu.device), reference[1].to(mu.device))
