"""Per-sample cost of one CIFAR NFE against the batch size (L2 residency vs wave quantisation)."""
import sys, time, torch, collections
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import __graft_entry__ as g
pkg = g.load_package()
from oracle import unet as O
from golden_configs import GOLDEN_CONFIGS
cfg = GOLDEN_CONFIGS["cifar"][0]
m = pkg.UNetModelWrapper(dim=(3, 32, 32), num_res_blocks=2, num_channels=128, channel_mult=[1, 2, 2, 2], num_heads=4,
                         num_head_channels=64, attention_resolutions="16", precision="bf16")
m.load_state_dict(O.seeded_params(cfg, 0)); m = m.cuda().eval()
e = m.engine()
out = open('gpurun_out/batch_sweep.txt', 'w')
for B in [int(a) for a in sys.argv[1:]] or [64, 128, 192, 256, 384, 512, 768, 1024]:
    x = torch.randn(B, 3, 32, 32, device='cuda')
    for _ in range(3): e.forward(x, 0.5)
    torch.cuda.synchronize()
    n = max(5, 4096 // B)
    t0 = time.time()
    for _ in range(n): e.forward(x, 0.5)
    torch.cuda.synchronize()
    dt = (time.time() - t0) / n
    rows = e.profile_forward(x, 0.5, repeats=5)
    agg = collections.defaultdict(float)
    for r in rows: agg[r['kind']] += r['ms']
    line = f"B={B:5d} NFE {dt*1e3:7.3f} ms  {dt*1e6/B:7.3f} us/sample  | " + "  ".join(f"{k} {v*1e3/B:6.3f}" for k, v in sorted(agg.items()))
    print(line); out.write(line + "\n")
