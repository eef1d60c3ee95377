"""Short NFE workload for ncu captures: CIFAR config, batch from argv (default 1024), 2 NFEs."""
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as g
g.build(); pkg = g.load_package()
from oracle import unet as O
from golden_configs import GOLDEN_CONFIGS
cfg = GOLDEN_CONFIGS["cifar"][0]
m = pkg.UNetModelWrapper(dim=(3, 32, 32), num_res_blocks=2, num_channels=128, channel_mult=[1, 2, 2, 2], num_heads=4,
                         num_head_channels=64, attention_resolutions="16", precision="bf16")
m.load_state_dict(O.seeded_params(cfg, 0)); m = m.cuda().eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
x = torch.randn(B, 3, 32, 32, device='cuda')
for _ in range(n):
    m.engine().forward(x, 0.5)
torch.cuda.synchronize()
print("done")
