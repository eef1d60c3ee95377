"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol that include/cfm_b200.h declares; no compute calls are made."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cfm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cfm_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ("cfm_engine_create", "cfm_engine_destroy", "cfm_engine_forward", "cfm_sample_euler",
              "cfm_sample_ddpm", "cfm_rk_combine", "cfm_rk_error_sumsq", "cfm_make_box_condition",
              "cfm_quantize_u8", "cfm_last_error", "cfm_ddpm_step", "cfm_sample_sde", "cfm_sde_em_step", "cfm_ddpm_em_step",
              "cfm_resize_bilinear", "cfm_fid_accumulate", "cfm_engine_op_info"):
        assert s in syms


def test_library_exports_every_declared_symbol(pkg):
    lib = ctypes.CDLL(pkg.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/cfm_b200.h but not exported"
    assert lib.cfm_abi_version() == 2


def test_python_binding_covers_the_header(pkg):
    assert sorted(pkg._lib.SIGNATURES) == declared_symbols()


def test_config_struct_layout_matches_header(pkg):
    # 6 ints + 8 floats + 1 int + 8 ints + 9 ints + flags + 6 reserved
    assert ctypes.sizeof(pkg._lib.UNetConfigC) == 4 * (6 + 8 + 1 + 8 + 9 + 7)
    assert ctypes.sizeof(pkg._lib.DdpmOptionsC) == 4 * 8


def test_create_fails_loudly_without_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from oracle import unet as O
    cfg = O.config_from_create_model(image_size=16, in_channels=3, out_channels=3, num_channels=32, num_res_blocks=1,
                                     channel_mult="1,2", attention_resolutions="8")
    m = pkg.UNetModel(image_size=16, in_channels=3, model_channels=32, out_channels=3, num_res_blocks=1,
                      attention_resolutions=(2,), channel_mult=(1, 2))
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 16, 16), torch.zeros(1))
    with pytest.raises(RuntimeError):
        pkg.Engine(m.config, m.state_dict(), device="cuda:0")


def test_null_arguments_are_rejected_not_crashed(pkg):
    lib = pkg._lib.load()
    assert lib.cfm_engine_forward(None, 1, None, None, None, 0.0, None, None, None) != 0
    assert lib.cfm_quantize_u8(None, None, 0, None) != 0
    assert lib.cfm_last_error(None) is not None
    assert lib.cfm_ddpm_step(None, None, None, None, None, 0, 0, None, 0, 0, None) != 0
    assert lib.cfm_resize_bilinear(None, None, 0, 1, 1, 1, 1, None) != 0
    assert lib.cfm_fid_accumulate(None, None, None, 0, 1, None) != 0
    assert lib.cfm_sample_sde(None, None, 1, None, None, None, None, 0, 0.1, None, 0, None) != 0
