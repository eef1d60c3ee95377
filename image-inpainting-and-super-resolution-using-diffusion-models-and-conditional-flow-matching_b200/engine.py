"""Python handle on a native engine (``cfm_engine*``): one U-Net on one GPU.

PyTorch is used only for device memory, streams and the tensors handed in/out; every
arithmetic op of an NFE runs in libcfm_b200.so.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib


@dataclasses.dataclass(frozen=True)
class UNetConfig:
    """Constructor surface of ``UNetModel`` (reference AD/image_diffusion/unet.py:518-539)."""
    image_size: int
    in_channels: int
    model_channels: int
    out_channels: int
    num_res_blocks: int
    attention_ds: Tuple[int, ...]
    channel_mult: Tuple[float, ...] = (1, 2, 4, 8)
    conv_resample: bool = True
    num_classes: Optional[int] = None
    num_heads: int = 1
    num_head_channels: int = -1
    num_heads_upsample: int = -1
    use_scale_shift_norm: bool = False
    resblock_updown: bool = False
    use_new_attention_order: bool = False
    fuse_groupnorm: bool = True      # engine option (not a reference hyper-parameter): fold out_layers GroupNorm into conv1

    def to_c(self, precision: int) -> _lib.UNetConfigC:
        if len(self.channel_mult) > _lib.MAX_LEVELS or len(self.attention_ds) > _lib.MAX_LEVELS:
            raise ValueError("too many levels")
        c = _lib.UNetConfigC()
        c.image_size, c.in_channels, c.model_channels = self.image_size, self.in_channels, self.model_channels
        c.out_channels, c.num_res_blocks, c.n_levels = self.out_channels, self.num_res_blocks, len(self.channel_mult)
        for i, m in enumerate(self.channel_mult):
            c.channel_mult[i] = float(m)
        c.n_attention_ds = len(self.attention_ds)
        for i, d in enumerate(self.attention_ds):
            c.attention_ds[i] = int(d)
        c.conv_resample = int(self.conv_resample)
        c.num_classes = int(self.num_classes or 0)
        c.num_heads, c.num_head_channels, c.num_heads_upsample = self.num_heads, self.num_head_channels, self.num_heads_upsample
        c.use_scale_shift_norm = int(self.use_scale_shift_norm)
        c.resblock_updown = int(self.resblock_updown)
        c.use_new_attention_order = int(self.use_new_attention_order)
        c.precision = precision
        c.flags = 0 if self.fuse_groupnorm else _lib.FLAG_SEPARATE_GROUPNORM
        return c


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _as_f32_cuda(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.float32).contiguous()


class Engine:
    """Owns a ``cfm_engine*``.  ``precision`` is "fp32" (exact mode) or "bf16" (tensor cores)."""

    def __init__(self, config: UNetConfig, state_dict: Dict[str, torch.Tensor], device=None, precision: str = "bf16"):
        if not torch.cuda.is_available():
            raise _lib.EngineError("no CUDA device: the sampling engine has no CPU fallback")
        self.lib = _lib.load()
        self.config = config
        self.precision = precision
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise _lib.EngineError("engine device must be a CUDA device")
        prec = {"fp32": _lib.PRECISION_FP32, "bf16": _lib.PRECISION_BF16}[precision]
        cfg = config.to_c(prec)
        items = [(k, v.detach().to("cpu", torch.float32).contiguous()) for k, v in state_dict.items()
                 if torch.is_tensor(v) and v.is_floating_point()]
        n = len(items)
        names = (C.c_char_p * n)(*[k.encode() for k, _ in items])
        ptrs = (C.c_void_p * n)(*[v.data_ptr() for _, v in items])
        numel = (C.c_int64 * n)(*[v.numel() for _, v in items])
        handle = C.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        with torch.cuda.device(self.device):      # cfm_engine_create selects the device: keep the caller's current one
            rc = self.lib.cfm_engine_create(C.byref(cfg), n, names, ptrs, numel, idx, C.byref(handle))
        _lib.check(rc, None)
        self._h = handle
        self.x_channels = config.out_channels
        self.cond_channels = config.in_channels - config.out_channels

    def close(self):
        if getattr(self, "_h", None):
            with torch.cuda.device(self.device):
                self.lib.cfm_engine_destroy(self._h)
            self._h = None

    def _labels(self, y: Optional[torch.Tensor], B: int) -> Optional[torch.Tensor]:
        """int64 labels on the device, range-checked (an out-of-range label would index past the embedding table)."""
        assert (y is not None) == (self.config.num_classes is not None), \
            "must specify y if and only if the model is class-conditional"
        if y is None:
            return None
        if tuple(y.shape) != (B,):
            raise ValueError(f"y must have shape ({B},), got {tuple(y.shape)}")
        yd = y.to(device=self.device, dtype=torch.int64).contiguous()
        if B and (int(yd.min()) < 0 or int(yd.max()) >= self.config.num_classes):
            raise IndexError(f"class labels must lie in [0, {self.config.num_classes})")
        return yd

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- introspection -------------------------------------------------------------------------
    @property
    def param_count(self) -> int:
        return int(self.lib.cfm_engine_param_count(self._h))

    @property
    def flops_per_sample(self) -> float:
        return float(self.lib.cfm_engine_flops_per_sample(self._h))

    @property
    def last_launches(self) -> int:
        return int(self.lib.cfm_engine_kernel_launches(self._h))

    @property
    def tensor_core_convs(self) -> int:
        return int(self.lib.cfm_engine_tensor_core_convs(self._h))

    @property
    def cached_graphs(self) -> int:
        return int(self.lib.cfm_engine_cached_graphs(self._h))

    def workspace_bytes(self, batch: int) -> int:
        return int(self.lib.cfm_engine_workspace_bytes(self._h, batch))

    # --- one NFE -------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, t, y: Optional[torch.Tensor] = None, cond: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [B,Cx,H,W]; t python float / 0-dim tensor (shared) or [B] tensor; y [B] int64; cond [B,Cc,H,W]."""
        S = self.config.image_size
        B = x.shape[0]
        cx = self.config.in_channels if cond is None else self.x_channels
        if tuple(x.shape[1:]) != (cx, S, S):
            raise ValueError(f"x must be [B,{cx},{S},{S}], got {tuple(x.shape)}")
        if cond is not None and tuple(cond.shape) != (B, self.config.in_channels - cx, S, S):
            raise ValueError(f"cond must be [B,{self.config.in_channels - cx},{S},{S}], got {tuple(cond.shape)}")
        xd = _as_f32_cuda(x, self.device)
        cd = None if cond is None else _as_f32_cuda(cond, self.device)
        yd = self._labels(y, B)
        t_dev, t_scalar = None, 0.0
        if torch.is_tensor(t):
            while t.dim() > 1:
                t = t[:, 0]
            if t.dim() == 0:
                t_scalar = float(t)
            else:
                assert t.shape[0] == B
                t_dev = _as_f32_cuda(t, self.device)
        else:
            t_scalar = float(t)
        if out is None:
            out = torch.empty((B, self.config.out_channels, S, S), device=self.device, dtype=torch.float32)
        if B == 0:          # empty batch: like the PyTorch module, an empty result (the C ABI rejects batch <= 0)
            return out
        with torch.cuda.device(self.device):
            rc = self.lib.cfm_engine_forward(self._h, B, _ptr(xd), _ptr(cd), _ptr(t_dev), t_scalar, _ptr(yd),
                                             _ptr(out), _stream_ptr(self.device))
        _lib.check(rc, self._h)
        return out

    def profile_forward(self, x: torch.Tensor, t: float = 0.5, y=None, cond=None, repeats: int = 3):
        """Per-op device times of one NFE: list of dicts {name, kind, ms, flops, flops_executed, bytes}, the last three
        for the whole batch: algorithmic 2*MAC, the 2*MAC the kernel issues, algorithmic bytes (operands + result once)."""
        B = x.shape[0]
        xd = _as_f32_cuda(x, self.device)
        cd = None if cond is None else _as_f32_cuda(cond, self.device)
        yd = None if y is None else y.to(device=self.device, dtype=torch.int64).contiguous()
        out = torch.empty((B, self.config.out_channels, self.config.image_size, self.config.image_size),
                          device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            rc = self.lib.cfm_engine_profile_forward(self._h, B, _ptr(xd), _ptr(cd), float(t), _ptr(yd), _ptr(out),
                                                     repeats, _stream_ptr(self.device))
        _lib.check(rc, self._h)
        kinds = {0: "conv_generic", 1: "groupnorm", 2: "resample", 3: "attention_generic", 4: "conv_tcgen05", 5: "attention"}
        rows = []
        for i in range(self.lib.cfm_engine_profile_count(self._h)):
            name = C.create_string_buffer(128)
            kind, ms, fl = C.c_int32(), C.c_double(), C.c_double()
            _lib.check(self.lib.cfm_engine_profile_get(self._h, i, name, 128, C.byref(kind), C.byref(ms), C.byref(fl)), self._h)
            ex, by = C.c_double(), C.c_double()
            _lib.check(self.lib.cfm_engine_op_info(self._h, i, 0, C.byref(ex)), self._h)
            _lib.check(self.lib.cfm_engine_op_info(self._h, i, 1, C.byref(by)), self._h)
            rows.append({"name": name.value.decode(), "kind": kinds[kind.value], "ms": ms.value, "flops": fl.value * B,
                         "flops_executed": ex.value * B, "bytes": by.value * B})
        return rows

    # --- fused fixed-step Euler loop ----------------------------------------------------------------
    def sample_euler(self, x0: torch.Tensor, t_grid: Sequence[float], dt_grid: Sequence[float],
                     y: Optional[torch.Tensor] = None, cond: Optional[torch.Tensor] = None,
                     cond_drift: bool = False, return_trajectory: bool = False, return_uint8: bool = False,
                     use_graph: bool = False, guidance_weight: Optional[float] = None):
        """Runs ``x += dt_k * model(t_k, x)`` for every k on the device.  Returns (x_final, traj|None, u8|None).

        ``guidance_weight`` (class-conditional models): classifier-free guidance, two evaluations per step,
        ``v = v_c + w (v_c - v_u)`` with ``v_u`` evaluated without the label embedding."""
        n_steps = len(t_grid)
        assert len(dt_grid) == n_steps
        B = x0.shape[0]
        S = self.config.image_size
        cx = self.config.in_channels if cond is None else self.x_channels
        if tuple(x0.shape[1:]) != (cx, S, S):
            raise ValueError(f"x0 must be [B,{cx},{S},{S}], got {tuple(x0.shape)}")
        if cond is not None and tuple(cond.shape) != (B, self.config.in_channels - cx, S, S):
            raise ValueError(f"cond must be [B,{self.config.in_channels - cx},{S},{S}], got {tuple(cond.shape)}")
        x = _as_f32_cuda(x0, self.device).clone()        # the native loop integrates in place: keep the caller's x0
        cd = None if cond is None else _as_f32_cuda(cond, self.device)
        if cond_drift and cd is not None and cd.data_ptr() == cond.data_ptr():
            cd = cd.clone()                               # COND_DRIFT writes the drifted conditioning back
        yd = self._labels(y, B)
        traj = torch.empty((n_steps + 1,) + tuple(x.shape), device=self.device, dtype=torch.float32) if return_trajectory else None
        img = torch.empty(tuple(x.shape), device=self.device, dtype=torch.uint8) if return_uint8 else None
        tg = (C.c_float * max(n_steps, 1))(*[float(v) for v in t_grid])
        dg = (C.c_float * max(n_steps, 1))(*[float(v) for v in dt_grid])
        flags = (_lib.EULER_COND_DRIFT if cond_drift else 0) | (_lib.EULER_USE_GRAPH if use_graph else 0)
        if B == 0:          # empty shard (more ranks than samples): nothing to integrate
            return x, traj, img
        with torch.cuda.device(self.device):
            if guidance_weight is None:
                rc = self.lib.cfm_sample_euler(self._h, B, _ptr(x), _ptr(cd), _ptr(yd), tg, dg, n_steps, flags,
                                               _ptr(traj), _ptr(img), _stream_ptr(self.device))
            else:
                rc = self.lib.cfm_sample_euler_cfg(self._h, B, _ptr(x), _ptr(cd), _ptr(yd), float(guidance_weight), tg, dg,
                                                   n_steps, flags, _ptr(traj), _ptr(img), _stream_ptr(self.device))
        _lib.check(rc, self._h)
        return x, traj, img

    # --- fused DDPM reverse chain -------------------------------------------------------------------
    def sample_ddpm(self, xT: torch.Tensor, tables: Dict[str, torch.Tensor], mode: str = "prior",
                    condition: Optional[torch.Tensor] = None, pad_value: float = -2.0,
                    replace_below_step: Optional[int] = None, noise_condition: bool = True,
                    noise: Optional[torch.Tensor] = None, seed: int = 0, use_graph: bool = False,
                    n_corrector: int = 0, corrector_delta: float = 0.1) -> torch.Tensor:
        Ns = int(tables["sqrt_alphas_cumprod"].numel())
        S = self.config.image_size
        if tuple(xT.shape[1:]) != (self.x_channels, S, S):
            raise ValueError(f"xT must be [B,{self.x_channels},{S},{S}], got {tuple(xT.shape)}")
        if mode != "prior" and condition is None:
            raise ValueError("conditional sampling needs a condition")
        if condition is not None:
            # the native chain reads batch * C * S * S floats of it: a broadcastable or differently shaped condition
            # (which torch.where / concat would accept or reject later) must match xT exactly here
            if condition.dim() == xT.dim() and condition.shape[0] == 1 and tuple(condition.shape[1:]) == tuple(xT.shape[1:]):
                condition = condition.expand_as(xT)
            if tuple(condition.shape) != tuple(xT.shape):
                raise ValueError(f"condition must have the shape of xT {tuple(xT.shape)}, got {tuple(condition.shape)}")
            if mode == "amortized" and self.cond_channels != self.x_channels:
                raise ValueError("amortized conditioning needs a network with 2 * C input channels")
        x = _as_f32_cuda(xT, self.device).clone()
        cd = None if condition is None else _as_f32_cuda(condition, self.device)
        keep = []
        tb = _lib.DdpmTablesC()
        tb.Ns = Ns
        for name in ("sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
                     "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2",
                     "posterior_log_variance_clipped", "model_time"):
            h = tables[name].detach().to("cpu", torch.float32).contiguous()
            assert h.numel() == Ns, name
            keep.append(h)
            setattr(tb, name, C.cast(h.data_ptr(), C.POINTER(C.c_float)))
        opt = _lib.DdpmOptionsC()
        opt.mode = {"prior": _lib.DDPM_PRIOR, "replacement": _lib.DDPM_REPLACEMENT, "amortized": _lib.DDPM_AMORTIZED}[mode]
        opt.pad_value = float(pad_value)
        opt.replace_below_step = Ns if replace_below_step is None else int(replace_below_step)
        opt.noise_condition = int(noise_condition)
        opt.use_graph = int(use_graph)
        opt.n_corrector = int(n_corrector)
        opt.corrector_delta = float(corrector_delta)
        nd = None
        if noise is not None:
            nd = _as_f32_cuda(noise, self.device)
            assert nd.numel() == Ns * (2 + int(n_corrector)) * x.numel(), "noise must be [Ns, 2 + n_corrector, B*C*H*W]"
        if x.shape[0] == 0:
            return x
        with torch.cuda.device(self.device):
            rc = self.lib.cfm_sample_ddpm(self._h, x.shape[0], _ptr(x), _ptr(cd), C.byref(tb), C.byref(opt),
                                          _ptr(nd), C.c_uint64(seed), _stream_ptr(self.device))
        _lib.check(rc, self._h)
        del keep
        return x
