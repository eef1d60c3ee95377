"""In-loop NFE time of the CIFAR sampler (power-capped steady state, CUDA graph) for same-box A/B of library builds:
    CFM_B200_LIB=variants/libA.so python profiles/ab_loop.py [B] [nfe] [repeats]"""
import sys, time, torch
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
nfe = int(sys.argv[2]) if len(sys.argv) > 2 else 100
rep = int(sys.argv[3]) if len(sys.argv) > 3 else 2
x = torch.randn(B, 3, 32, 32, device='cuda')
t_span = torch.linspace(0, 1, nfe + 1)
pkg.sample_euler(m, x, t_span, return_uint8=True, use_graph=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(rep):
    pkg.sample_euler(m, x, t_span, return_uint8=True, use_graph=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / rep / nfe
print(f"lib={pkg.LIB_PATH.split('/')[-1]} B={B} ms/NFE={ms:.3f} samples/s={B / (ms * nfe) * 1e3:.1f} launches/NFE={m.engine().last_launches / nfe:.0f}")
