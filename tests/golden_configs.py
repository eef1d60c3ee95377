"""Configurations shared by the golden-vector generator and the parity tests."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import unet as O  # noqa: E402

# name -> (oracle UNetConfig, batch, weight seed)
GOLDEN_CONFIGS = {
    # BASELINE config 2: CIFAR-10 CFM (cifar10/compute_fid.py:39-48)
    "cifar": (O.config_from_wrapper((3, 32, 32), 128, 2, channel_mult=[1, 2, 2, 2], num_heads=4,
                                    num_head_channels=64, attention_resolutions="16"), 2, 0),
    # BASELINE config 1/3: MNIST CFM wrapper defaults (mnist/train_mnist.py:256; conditional_mnist.ipynb)
    "mnist_cfm": (O.config_from_wrapper((1, 28, 28), 32, 1), 2, 1),
    # InPaintModelWrapper layout (2 input channels, 1 output)
    "mnist_inpaint": (O.config_from_wrapper((1, 28, 28), 32, 1, extra_in_channels=1), 2, 2),
    # BASELINE config 4: DDPM MNIST, amortised (AD/experiments/config.py:101-107; 2-ch in)
    "mnist_ddpm": (O.config_from_create_model(image_size=28, in_channels=2, out_channels=1, num_channels=32,
                                              num_res_blocks=1, channel_mult="1, 2, 2", resblock_updown=True), 2, 3),
    # BASELINE config 4: DDPM Flowers-64 (AD/experiments/config.py:108-116)
    "flowers_ddpm": (O.config_from_create_model(image_size=64, in_channels=3, out_channels=3, num_channels=128,
                                                num_res_blocks=1, resblock_updown=True, num_head_channels=64,
                                                use_scale_shift_norm=True, num_heads=4), 1, 4),
    # small odd-ball: new attention order, 2 heads, no conv_resample path is not constructible via create_model
    "tiny_neworder": (O.config_from_create_model(image_size=16, in_channels=3, out_channels=3, num_channels=32,
                                                 num_res_blocks=1, channel_mult="1,2", attention_resolutions="8",
                                                 num_heads=2, use_new_attention_order=True), 3, 5),
}

# --- whole DDPM reverse chains, run through the reference's sampling.py (tests/golden/ddpm_chains.npz) -----------------
CHAIN_NS = 24      # beta_max = 20 / Ns must stay below 1 (sde_diffusion.py:14-15)


def chain_net_cfg(in_ch):
    return O.config_from_create_model(image_size=16, in_channels=in_ch, out_channels=1, num_channels=32, num_res_blocks=1,
                                      channel_mult="1,2", attention_resolutions="8", resblock_updown=True)


def chain_inputs(seed):
    """xT and an in-painting condition (6x6 hole := -2) from a frozen numpy stream."""
    import numpy as np
    import torch
    rs = np.random.RandomState(500 + seed)
    xT = torch.from_numpy(rs.standard_normal((2, 1, 16, 16)).astype(np.float32))
    cond = torch.from_numpy(rs.uniform(-1, 1, size=(2, 1, 16, 16)).astype(np.float32))
    cond[0, :, 5:11, 6:12] = -2.0
    cond[1, :, 7:13, 5:11] = -2.0
    return xT, cond


def _case(kind, in_ch, seed, prior=False, n_corrector=0, delta=0.1, start_fraction=1.0, noise=True):
    return dict(kind=kind, in_ch=in_ch, seed=seed, prior=prior, n_corrector=n_corrector, delta=delta,
                start_fraction=start_fraction, noise=noise)


CHAIN_CASES = {
    "prior": _case("replacement", 1, 101, prior=True),
    "prior_amortized": _case("amortized", 2, 102, prior=True),
    "replacement": _case("replacement", 1, 103, start_fraction=0.75),
    "replacement_raw_condition": _case("replacement", 1, 104, noise=False),
    "replacement_corrector2": _case("replacement", 1, 105, n_corrector=2, delta=0.1),
    "amortized": _case("amortized", 2, 106),
    "amortized_corrector1": _case("amortized", 2, 107, n_corrector=1, delta=0.2),
}


# --- seeded sweep over the constructor's keyword space (tests/test_oracle.py pins the oracle on it against the live
# reference, tests/test_gpu_unet.py the engine against the oracle) ---------------------------------------------------
N_RANDOM_CONFIGS = 16


def random_config(seed):
    """A seeded draw from the reference's create_model keyword space (unet.py:43-105) inside the GroupNorm32 domain
    (every channel count a multiple of 32), small enough for the CPU oracle."""
    import numpy as np
    rs = np.random.RandomState(7000 + seed)
    size = int(rs.choice([8, 12, 16, 20, 32]))
    levels = int(rs.randint(1, 4 if size % 8 == 0 else 3))
    while size % (1 << (levels - 1)):
        levels -= 1
    mult = tuple(int(rs.choice([1, 2, 3])) for _ in range(levels))
    nc = int(rs.choice([32, 64, 96]))
    attn_levels = [lv for lv in range(levels) if rs.rand() < 0.6 and (size >> lv) <= 16]
    kw = dict(image_size=size, in_channels=int(rs.choice([1, 2, 3, 6])), out_channels=int(rs.choice([1, 3])), num_channels=nc,
              num_res_blocks=int(rs.randint(1, 3)), channel_mult=",".join(map(str, mult)),
              attention_resolutions=",".join(str(size >> lv) for lv in attn_levels) if attn_levels else str(4 * size),
              use_scale_shift_norm=bool(rs.rand() < 0.35), resblock_updown=bool(rs.rand() < 0.4),
              use_new_attention_order=bool(rs.rand() < 0.4))
    widths = {nc * mult[lv] for lv in attn_levels} | {nc * mult[-1]}        # channels of every attention block (middle included)
    hcs = [h for h in (32, 64) if all(w % h == 0 for w in widths)]
    if rs.rand() < 0.5 and hcs:
        kw["num_head_channels"] = int(rs.choice(hcs))
    else:
        kw["num_heads"] = int(rs.choice([h for h in (1, 2, 4) if all(w % h == 0 for w in widths)]))
        if rs.rand() < 0.3:
            kw["num_heads_upsample"] = int(rs.choice([1, 2]))
    if rs.rand() < 0.3:
        kw["num_classes"] = int(rs.randint(2, 12))
    return kw, int(rs.randint(1, 8))
