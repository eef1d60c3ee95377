"""Short NFE workload for ncu captures: class-conditional MNIST config (28x28, T = 784 attention), batch from argv."""
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as g
g.build(); pkg = g.load_package()
from oracle import unet as O
cfg = O.config_from_wrapper((1, 28, 28), 32, 1, class_cond=True, num_classes=10)
m = pkg.UNetModelWrapper(dim=(1, 28, 28), num_channels=32, num_res_blocks=1, num_classes=10, class_cond=True, precision="bf16")
m.load_state_dict(O.seeded_params(cfg, 0)); m = m.cuda().eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
x = torch.randn(B, 1, 28, 28, device='cuda'); y = torch.arange(B, device='cuda') % 10
for _ in range(n):
    m.engine().forward(x, 0.5, y=y)
torch.cuda.synchronize()
print("done")
