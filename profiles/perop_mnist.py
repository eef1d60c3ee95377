"""Per-op timing of one NFE of the class-conditional MNIST config (batch from argv, default 4096)."""
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as g
g.build(); pkg = g.load_package()
from oracle import unet as O
cfg = O.config_from_wrapper((1, 28, 28), 32, 1, class_cond=True, num_classes=10)
m = pkg.UNetModelWrapper(dim=(1, 28, 28), num_channels=32, num_res_blocks=1, num_classes=10, class_cond=True, precision="bf16")
m.load_state_dict(O.seeded_params(cfg, 0)); m = m.cuda().eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x = torch.randn(B, 1, 28, 28, device='cuda'); y = torch.arange(B, device='cuda') % 10
e = m.engine()
rows = e.profile_forward(x, 0.5, y=y, repeats=5)
tot = sum(r['ms'] for r in rows)
print(f"total {tot:.3f} ms")
for r in rows:
    print(f"{r['name']:36s} {r['kind']:13s} {r['ms']:8.4f} ms {r['flops']/r['ms']/1e9 if r['ms'] else 0:8.1f} TF/s")
