"""Short sampler workloads for ncu captures of the step kernels: 3 Euler steps (CIFAR, batch from argv, no graph so every
launch is visible) and 3 DDPM steps of a Replacement chain on the same network shape."""
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as g
pkg = g.load_package()
from oracle import unet as O
from golden_configs import GOLDEN_CONFIGS
cfg = GOLDEN_CONFIGS["cifar"][0]
m = pkg.UNetModelWrapper(dim=(3, 32, 32), num_res_blocks=2, num_channels=128, channel_mult=[1, 2, 2, 2], num_heads=4,
                         num_head_channels=64, attention_resolutions="16", precision="bf16")
m.load_state_dict(O.seeded_params(cfg, 0)); m = m.cuda().eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x = torch.randn(B, 3, 32, 32, device='cuda')
pkg.sample_euler(m, x, torch.linspace(0, 1, 4), return_uint8=True, use_graph=False)
ddpm = pkg.DDPM(3)
cond = torch.rand(B, 3, 32, 32, device='cuda') * 2 - 1
cond[:, :, 8:20, 8:20] = -2.0
m.engine().sample_ddpm(x, ddpm.tables(), mode="replacement", condition=cond, seed=1)
torch.cuda.synchronize()
print("done")
