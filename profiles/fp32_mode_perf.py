import sys, time, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as g
pkg = g.load_package()
from oracle import unet as O
from golden_configs import GOLDEN_CONFIGS
cfg = GOLDEN_CONFIGS["cifar"][0]
m = pkg.UNetModelWrapper(dim=(3, 32, 32), num_res_blocks=2, num_channels=128, channel_mult=[1, 2, 2, 2], num_heads=4,
                         num_head_channels=64, attention_resolutions="16", precision="fp32")
m.load_state_dict(O.seeded_params(cfg, 0)); m = m.cuda().eval()
for B in (16, 128):
    x = torch.randn(B, 3, 32, 32, device='cuda')
    e = m.engine()
    e.forward(x, 0.5); torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3): e.forward(x, 0.5)
    torch.cuda.synchronize(); dt = (time.time() - t0) / 3
    print(f"fp32 exact mode B={B}: NFE {dt*1e3:.2f} ms -> {e.flops_per_sample*B/dt/1e12:.1f} TFLOP/s, {B/(dt*100):.1f} samples/s at 100 NFE")
    rows = e.profile_forward(x, 0.5, repeats=2)
    import collections
    agg = collections.defaultdict(float)
    for r in rows: agg[r['kind']] += r['ms']
    print("   ", dict(agg))
