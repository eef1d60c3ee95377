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
