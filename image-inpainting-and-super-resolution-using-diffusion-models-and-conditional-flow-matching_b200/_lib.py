"""ctypes binding of libcfm_b200.so (the C ABI declared in include/cfm_b200.h).

The library is built in-tree by ``__graft_entry__.build()``.  There is no CPU or
PyTorch fallback: a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CFM_B200_LIB") or os.path.join(_HERE, "libcfm_b200.so")   # env override: A/B kernel variants
MAX_LEVELS = 8

PRECISION_FP32 = 0
PRECISION_BF16 = 1
FLAG_SEPARATE_GROUPNORM = 1
EULER_COND_DRIFT = 1
EULER_USE_GRAPH = 2
DDPM_PRIOR, DDPM_REPLACEMENT, DDPM_AMORTIZED = 0, 1, 2


class UNetConfigC(C.Structure):
    _fields_ = [
        ("image_size", C.c_int32), ("in_channels", C.c_int32), ("model_channels", C.c_int32),
        ("out_channels", C.c_int32), ("num_res_blocks", C.c_int32), ("n_levels", C.c_int32),
        ("channel_mult", C.c_float * MAX_LEVELS), ("n_attention_ds", C.c_int32),
        ("attention_ds", C.c_int32 * MAX_LEVELS), ("conv_resample", C.c_int32),
        ("num_classes", C.c_int32), ("num_heads", C.c_int32), ("num_head_channels", C.c_int32),
        ("num_heads_upsample", C.c_int32), ("use_scale_shift_norm", C.c_int32),
        ("resblock_updown", C.c_int32), ("use_new_attention_order", C.c_int32),
        ("precision", C.c_int32), ("flags", C.c_int32), ("reserved", C.c_int32 * 6),
    ]


class DdpmTablesC(C.Structure):
    _fields_ = [("Ns", C.c_int32)] + [(n, C.POINTER(C.c_float)) for n in (
        "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
        "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2",
        "posterior_log_variance_clipped", "model_time")]


class DdpmOptionsC(C.Structure):
    _fields_ = [("mode", C.c_int32), ("pad_value", C.c_float), ("replace_below_step", C.c_int32),
                ("noise_condition", C.c_int32), ("use_graph", C.c_uint32), ("n_corrector", C.c_uint32),
                ("corrector_delta", C.c_float), ("reserved", C.c_uint32 * 1)]


# name -> (restype, argtypes); mirrors include/cfm_b200.h one to one
SIGNATURES = {
    "cfm_abi_version": (C.c_int, []),
    "cfm_last_error": (C.c_char_p, [C.c_void_p]),
    "cfm_engine_create": (C.c_int, [C.POINTER(UNetConfigC), C.c_int32, C.POINTER(C.c_char_p),
                                    C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int32,
                                    C.POINTER(C.c_void_p)]),
    "cfm_engine_destroy": (None, [C.c_void_p]),
    "cfm_engine_param_count": (C.c_int64, [C.c_void_p]),
    "cfm_engine_flops_per_sample": (C.c_double, [C.c_void_p]),
    "cfm_engine_workspace_bytes": (C.c_int64, [C.c_void_p, C.c_int32]),
    "cfm_engine_kernel_launches": (C.c_int32, [C.c_void_p]),
    "cfm_engine_tensor_core_convs": (C.c_int32, [C.c_void_p]),
    "cfm_engine_cached_graphs": (C.c_int32, [C.c_void_p]),
    "cfm_engine_forward": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cfm_engine_profile_forward": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p,
                                             C.c_void_p, C.c_int32, C.c_void_p]),
    "cfm_engine_profile_count": (C.c_int32, [C.c_void_p]),
    "cfm_engine_profile_get": (C.c_int, [C.c_void_p, C.c_int32, C.c_char_p, C.c_int32, C.POINTER(C.c_int32),
                                         C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "cfm_engine_op_info": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_double)]),
    "cfm_sample_euler": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int32, C.c_uint32,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "cfm_sample_euler_cfg": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                       C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int32, C.c_uint32,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "cfm_sample_ddpm": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(DdpmTablesC),
                                  C.POINTER(DdpmOptionsC), C.c_void_p, C.c_uint64, C.c_void_p]),
    "cfm_rk_combine": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_float),
                                 C.c_int32, C.c_float, C.c_int64, C.c_void_p]),
    "cfm_rk_error_sumsq": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_float), C.c_int32, C.c_float, C.c_float, C.c_float,
                                     C.c_int64, C.c_void_p]),
    "cfm_rk_scaled_sumsq": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int64,
                                      C.c_void_p]),
    "cfm_rk_dense_output": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                      C.c_float, C.c_int64, C.c_void_p]),
    "cfm_make_box_condition": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                         C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_void_p]),
    "cfm_quantize_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "cfm_ddpm_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(DdpmTablesC), C.POINTER(DdpmOptionsC),
                                C.c_int32, C.c_int32, C.c_void_p, C.c_uint64, C.c_int64, C.c_void_p]),
    "cfm_sde_em_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_uint64,
                                  C.c_uint32, C.c_int64, C.c_void_p]),
    "cfm_ddpm_em_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_double, C.c_void_p, C.c_uint64,
                                   C.c_uint32, C.c_int64, C.c_void_p]),
    "cfm_sample_sde": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_float),
                                 C.POINTER(C.c_float), C.c_int32, C.c_float, C.c_void_p, C.c_uint64, C.c_void_p]),
    "cfm_resize_bilinear": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_void_p]),
    "cfm_fid_accumulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
}

_lib = None


class EngineError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise EngineError(f"{LIB_PATH} is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback for the sampling engine)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        if os.environ.get("CFM_B200_LIB") and not hasattr(lib, name):
            continue                   # A/B runs against an older build of the library (profiles/build_variant.sh)
        fn = getattr(lib, name)        # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, handle=None):
    if rc != 0:
        msg = load().cfm_last_error(handle)
        raise EngineError(f"cfm_b200 error {rc}: {msg.decode() if msg else '?'}")
