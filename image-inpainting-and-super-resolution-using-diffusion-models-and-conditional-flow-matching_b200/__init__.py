"""B200-native sampling engine for CFM / DDPM image U-Nets (hot path of the reference repo).

Public surface mirrors the reference's seams for this path; everything numeric runs in
``libcfm_b200.so`` (hand-written sm_100a CUDA).  Importing this package does not need a GPU;
constructing an engine does, and fails loudly without one.
"""
from . import _lib
from ._lib import EngineError, LIB_PATH
from .engine import Engine, UNetConfig
from .models import (UNetModel, UNetModelWrapper, InPaintModelWrapper, SuperResModelWrapper, create_model,
                     load_checkpoint, parameter_layout, default_channel_mult)
from .integrators import (NeuralODE, odeint, sample_euler, sample_sde, euler_time_grid, rk_combine, rk_error_sumsq,
                          rk_scaled_sumsq, rk_dense_output)
from .diffusion import (DDPM, EpsModel, Amortized, Replacement, ReconstructionGuidance, InPainting, OutPainting,
                        HyperResolution, get_conditioning, get_likelihood, get_prior_sample_fn,
                        get_conditional_sample_fn, downsample_images, resize_bilinear, extract)
from .fid import FIDStatistics, frechet_distance, compute_fid
from .distributed import shard_range, sample_euler_sharded, gather_uint8, odeint_sharded

__all__ = [n for n in dir() if not n.startswith("_")]
