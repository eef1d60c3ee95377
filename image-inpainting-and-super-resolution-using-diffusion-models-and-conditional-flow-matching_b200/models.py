"""Drop-in model objects for the reference's vector-field / eps-network seams.

Same constructor keywords, ``state_dict`` key names and call signatures as the
reference's models, but ``forward`` runs on the native engine:

* ``UNetModel(image_size=..., ...)(x, timesteps)``   <- AD/image_diffusion/unet.py:490-728
* ``create_model(image_size=..., ...)``               <- AD/image_diffusion/unet.py:43-125
* ``UNetModelWrapper(dim=..., ...)(t, x, y=None)``    <- torchcfm wrapper; cifar10/compute_fid.py:39-48,70,83
* ``InPaintModelWrapper(...)(x, t, con=)``            <- mnist/utils_mnist.py:97, mnist/train_mnist.py:262-267
* ``SuperResModelWrapper(...)(x, t, low_res=)``       <- mnist/utils_mnist_hy.py:82, mnist/train_mnist_hy.py:312-317

The modules are ``nn.Module``s holding real ``nn.Parameter``s under the reference's names, so
``load_state_dict`` / ``state_dict`` / ``.to(device)`` / ``.eval()`` behave as users expect; the
native engine is (re)built lazily from the current parameter values on the first forward
after construction, ``load_state_dict`` or ``refresh()``.  Inference only: outputs carry no grad.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .engine import Engine, UNetConfig

NUM_CLASSES = 1000  # torchcfm's default


def default_channel_mult(image_size: int) -> Tuple[float, ...]:
    table = {512: (0.5, 1, 1, 2, 2, 4, 4), 256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4),
             32: (1, 2, 2, 2), 28: (1, 2, 2)}
    if image_size not in table:
        raise ValueError(f"unsupported image size: {image_size}")
    return table[image_size]


def _heads_for(cfg: UNetConfig, channels: int, upsample_side: bool) -> int:
    if cfg.num_head_channels != -1:
        assert channels % cfg.num_head_channels == 0, \
            f"q,k,v channels {channels} is not divisible by num_head_channels {cfg.num_head_channels}"
        return channels // cfg.num_head_channels
    if upsample_side and cfg.num_heads_upsample != -1:
        return cfg.num_heads_upsample
    return cfg.num_heads


def parameter_layout(cfg: UNetConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, init) in ``state_dict`` order.  init in {conv, zero_conv, linear, ones, zeros, embed}."""
    out: List[Tuple[str, Tuple[int, ...], str]] = []
    mc, ted = cfg.model_channels, 4 * cfg.model_channels

    def conv(p, co, ci, k, zero=False, one_d=False):
        shape = (co, ci, k) if one_d else (co, ci, k, k)
        out.append((f"{p}.weight", shape, "zero" if zero else "conv"))
        out.append((f"{p}.bias", (co,), "zero" if zero else "conv_bias"))

    def lin(p, co, ci):
        out.append((f"{p}.weight", (co, ci), "conv")); out.append((f"{p}.bias", (co,), "conv_bias"))

    def norm(p, c):
        out.append((f"{p}.weight", (c,), "ones")); out.append((f"{p}.bias", (c,), "zero"))

    def res(p, ci, co):
        norm(f"{p}.in_layers.0", ci); conv(f"{p}.in_layers.2", co, ci, 3)
        lin(f"{p}.emb_layers.1", 2 * co if cfg.use_scale_shift_norm else co, ted)
        norm(f"{p}.out_layers.0", co); conv(f"{p}.out_layers.3", co, co, 3, zero=True)
        if ci != co:
            conv(f"{p}.skip_connection", co, ci, 1)

    def attn(p, c):
        norm(f"{p}.norm", c); conv(f"{p}.qkv", 3 * c, c, 1, one_d=True); conv(f"{p}.proj_out", c, c, 1, zero=True, one_d=True)

    lin("time_embed.0", ted, mc); lin("time_embed.2", ted, ted)
    if cfg.num_classes is not None:
        out.append(("label_emb.weight", (cfg.num_classes, ted), "embed"))
    ch = int(cfg.channel_mult[0] * mc)
    conv("input_blocks.0.0", ch, cfg.in_channels, 3)
    chans = [ch]
    ds, idx = 1, 1
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            res(f"input_blocks.{idx}.0", ch, int(mult * mc)); ch = int(mult * mc)
            if ds in cfg.attention_ds:
                _heads_for(cfg, ch, False); attn(f"input_blocks.{idx}.1", ch)
            chans.append(ch); idx += 1
        if level != len(cfg.channel_mult) - 1:
            if cfg.resblock_updown:
                res(f"input_blocks.{idx}.0", ch, ch)
            elif cfg.conv_resample:
                conv(f"input_blocks.{idx}.0.op", ch, ch, 3)
            chans.append(ch); ds *= 2; idx += 1
    res("middle_block.0", ch, ch); attn("middle_block.1", ch); res("middle_block.2", ch, ch)
    idx = 0
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            ich = chans.pop()
            sub = 0
            res(f"output_blocks.{idx}.{sub}", ch + ich, int(mc * mult)); ch = int(mc * mult); sub += 1
            if ds in cfg.attention_ds:
                attn(f"output_blocks.{idx}.{sub}", ch); sub += 1
            if level and i == cfg.num_res_blocks:
                if cfg.resblock_updown:
                    res(f"output_blocks.{idx}.{sub}", ch, ch)
                elif cfg.conv_resample:
                    conv(f"output_blocks.{idx}.{sub}.conv", ch, ch, 3)
                ds //= 2
            idx += 1
    norm("out.0", ch)
    conv("out.2", cfg.out_channels, int(cfg.channel_mult[0] * mc), 3, zero=True)
    return out


class _Node(nn.Module):
    """Anonymous container so parameters sit at the reference's dotted paths."""


class UNetModel(nn.Module):
    """Engine-backed ``UNetModel``: ``forward(x, timesteps, y=None)`` (unet.py:708-728)."""

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None,
                 use_checkpoint=False, use_fp16=False, num_heads=1, num_head_channels=-1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=False,
                 precision: str = "bf16", fuse_groupnorm: bool = True):
        super().__init__()
        if dims != 2:
            raise NotImplementedError("the engine implements the 2-D U-Net only")
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        self.config = UNetConfig(image_size=image_size, in_channels=in_channels, model_channels=model_channels,
                                 out_channels=out_channels, num_res_blocks=num_res_blocks,
                                 attention_ds=tuple(attention_resolutions), channel_mult=tuple(channel_mult),
                                 conv_resample=conv_resample, num_classes=num_classes, num_heads=num_heads,
                                 num_head_channels=num_head_channels, num_heads_upsample=num_heads_upsample,
                                 use_scale_shift_norm=use_scale_shift_norm, resblock_updown=resblock_updown,
                                 use_new_attention_order=use_new_attention_order, fuse_groupnorm=fuse_groupnorm)
        self.image_size, self.in_channels, self.model_channels = image_size, in_channels, model_channels
        self.out_channels, self.num_classes, self.dropout = out_channels, num_classes, dropout
        self.precision = precision
        self._engine: Optional[Engine] = None
        self._engine_key = None
        for name, shape, init in parameter_layout(self.config):
            self._register(name, self._init_tensor(shape, init))

    @staticmethod
    def _init_tensor(shape, init) -> torch.Tensor:
        t = torch.empty(shape)
        if init == "zero":
            return t.zero_()
        if init == "ones":
            return t.fill_(1.0)
        if init == "embed":
            return t.normal_()
        fan_in = int(math.prod(shape[1:])) if len(shape) > 1 else None
        if init == "conv":
            return nn.init.kaiming_uniform_(t, a=math.sqrt(5))
        bound = 1.0 / math.sqrt(max(shape[0], 1)) if fan_in is None else 1.0 / math.sqrt(fan_in)
        return t.uniform_(-bound, bound)   # conv_bias (fan_in unknown here: harmless for inference drop-in)

    def _register(self, dotted: str, value: torch.Tensor):
        parts = dotted.split(".")
        node: nn.Module = self
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, _Node())
            node = node._modules[p]
        node.register_parameter(parts[-1], nn.Parameter(value, requires_grad=False))

    # -- engine lifecycle --------------------------------------------------------------------------
    def refresh(self):
        """Drop the native engine; it is rebuilt from the current parameters on the next forward."""
        if self._engine is not None:
            self._engine.close()
        self._engine = None

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        r = super().load_state_dict(state_dict, strict=strict, **kw)
        self.refresh()
        return r

    def _apply(self, fn, *a, **kw):
        r = super()._apply(fn, *a, **kw)
        self.refresh()
        return r

    def engine(self, device=None) -> Engine:
        p = next(self.parameters())
        dev = torch.device(device) if device is not None else p.device
        if dev.type != "cuda":
            raise RuntimeError("the B200 sampling engine needs its parameters on a CUDA device "
                               "(call .to('cuda')); there is no CPU fallback")
        key = (str(dev), self.precision)
        if self._engine is None or self._engine_key != key:
            self.refresh()
            self._engine = Engine(self.config, self.state_dict(), device=dev, precision=self.precision)
            self._engine_key = key
        return self._engine

    @torch.no_grad()
    def forward(self, x, timesteps, y=None):
        assert (y is not None) == (self.num_classes is not None), \
            "must specify y if and only if the model is class-conditional"
        return self.engine().forward(x, timesteps, y=y)


def create_model(*, image_size: int, in_channels: int, out_channels: int, num_channels: int, num_res_blocks,
                 channel_mult="", use_checkpoint=False, attention_resolutions="16", num_heads=1,
                 num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False, dropout=0,
                 resblock_updown=False, use_fp16=False, use_new_attention_order=False, model_path="",
                 precision: str = "bf16") -> UNetModel:
    """Keyword-compatible with the reference's ``create_model`` (unet.py:43-125), incl. checkpoint ingest."""
    if channel_mult == "":
        if image_size not in (512, 256, 128, 64):
            raise ValueError(f"unsupported image size: {image_size}")
        channel_mult = default_channel_mult(image_size)
    else:
        channel_mult = tuple(int(v) for v in channel_mult.split(","))
    if isinstance(attention_resolutions, int):
        attention_ds = [image_size // attention_resolutions]
    elif isinstance(attention_resolutions, str):
        attention_ds = [image_size // int(r) for r in attention_resolutions.split(",")]
    else:
        raise NotImplementedError
    model = UNetModel(image_size=image_size, in_channels=in_channels, model_channels=num_channels,
                      out_channels=out_channels, num_res_blocks=num_res_blocks,
                      attention_resolutions=tuple(attention_ds), dropout=dropout, channel_mult=channel_mult,
                      num_classes=None, num_heads=num_heads, num_head_channels=num_head_channels,
                      num_heads_upsample=num_heads_upsample, use_scale_shift_norm=use_scale_shift_norm,
                      resblock_updown=resblock_updown, use_new_attention_order=use_new_attention_order,
                      precision=precision)
    if model_path:
        load_checkpoint(model, model_path)
    return model


def load_checkpoint(model: nn.Module, path_or_state, strict: bool = False):
    """Checkpoint ingest for both reference formats.

    * DDPM: ``{"ema": {"ema_model.<name>": tensor}}`` (unet.py:107-115)
    * torchcfm: ``{"ema_model": state_dict}`` with an optional 7-char ``module.`` prefix
      (cifar10/compute_fid.py:54-64)
    """
    state = torch.load(path_or_state, map_location="cpu") if isinstance(path_or_state, (str, bytes)) else path_or_state
    if "ema" in state and isinstance(state["ema"], dict):
        state = {k[len("ema_model."):]: v for k, v in state["ema"].items() if "ema_model" in str(k)}
    elif "ema_model" in state and isinstance(state["ema_model"], dict):
        state = state["ema_model"]
    if state and all(k.startswith("module.") for k in state):
        state = {k[7:]: v for k, v in state.items()}
    return model.load_state_dict(state, strict=strict)


def _expand_t(t, batch: int):
    if not torch.is_tensor(t):
        return float(t)
    while t.dim() > 1:
        t = t[:, 0]
    return t


class UNetModelWrapper(UNetModel):
    """torchcfm's ``UNetModelWrapper(dim=(C,H,W), ...)``; called as ``model(t, x, y=None, *args, **kwargs)``.

    ``forward`` keeps a parameter literally named ``t`` (torchdyn inspects the signature) and
    tolerates torchdyn's extra ``args=`` keyword.
    """

    def __init__(self, dim, num_channels, num_res_blocks, channel_mult=None, learn_sigma=False, class_cond=False,
                 num_classes=NUM_CLASSES, use_checkpoint=False, attention_resolutions="16", num_heads=1,
                 num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False, dropout=0,
                 resblock_updown=False, use_fp16=False, use_new_attention_order=False, precision: str = "bf16",
                 _extra_in_channels: int = 0):
        image_size = dim[-1]
        if channel_mult is None:
            channel_mult = default_channel_mult(image_size)
        elif isinstance(channel_mult, str):
            channel_mult = tuple(int(v) for v in channel_mult.split(","))
        attention_ds = [image_size // int(r) for r in str(attention_resolutions).split(",")]
        super().__init__(image_size=image_size, in_channels=dim[0] + _extra_in_channels, model_channels=num_channels,
                         out_channels=(dim[0] if not learn_sigma else dim[0] * 2), num_res_blocks=num_res_blocks,
                         attention_resolutions=tuple(attention_ds), dropout=dropout, channel_mult=tuple(channel_mult),
                         num_classes=(num_classes if class_cond else None), num_heads=num_heads,
                         num_head_channels=num_head_channels, num_heads_upsample=num_heads_upsample,
                         use_scale_shift_norm=use_scale_shift_norm, resblock_updown=resblock_updown,
                         use_new_attention_order=use_new_attention_order, precision=precision)

    @torch.no_grad()
    def forward(self, t, x, y=None, *args, **kwargs):
        assert (y is not None) == (self.num_classes is not None), \
            "must specify y if and only if the model is class-conditional"
        return self.engine().forward(x, _expand_t(t, x.shape[0]), y=y)


class InPaintModelWrapper(UNetModelWrapper):
    """``model(x, t, con)``: conditioning image (pad_value holes) concatenated on channels.

    The reference's source for this wrapper is lost; only the call sites
    survive.  NOTE the (x, t) argument order.
    """

    def __init__(self, dim, num_channels, num_res_blocks, **kw):
        # call sites pass num_classes=None, class_cond=True -> not class conditional
        super().__init__(dim, num_channels, num_res_blocks, _extra_in_channels=dim[0], **kw)

    @torch.no_grad()
    def forward(self, x, t, con=None, *args, **kwargs):
        assert con is not None, "InPaintModelWrapper needs con="
        return self.engine().forward(x, _expand_t(t, x.shape[0]), cond=con)


class SuperResModelWrapper(InPaintModelWrapper):
    """``model(x, t, low_res=)``: low-res image bilinearly upsampled to x's size, concatenated on channels."""

    @torch.no_grad()
    def forward(self, x, t, low_res=None, *args, **kwargs):
        assert low_res is not None, "SuperResModelWrapper needs low_res="
        from .diffusion import resize_bilinear
        up = resize_bilinear(low_res, tuple(x.shape[-2:]))
        return self.engine().forward(x, _expand_t(t, x.shape[0]), cond=up)
