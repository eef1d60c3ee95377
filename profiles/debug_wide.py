"""Debug helper: one odd-shape config (index argv[1]) in bf16 against the oracle."""
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as g
pkg = g.load_package()
from oracle import unet as O
CFGS = [(16, (2, 4, 6), 1, -1, "1,2"), (8, (3, 6), -1, 192, "1"), (8, (8,), 1, -1, "1"), (16, (2,), 1, -1, "1"), (8, (4,), 1, -1, "1"), (4, (6,), 1, -1, "1")]
size, mult, heads, hc, attn = CFGS[int(sys.argv[1])]
kw = dict(channel_mult=list(mult), attention_resolutions=",".join(str(size // int(a)) for a in attn.split(",")))
if hc > 0: kw.update(num_head_channels=hc)
else: kw.update(num_heads=heads)
cfg = O.config_from_wrapper((3, size, size), 64, 1, **kw)
params = O.seeded_params(cfg, 31)
m = pkg.UNetModelWrapper(dim=(3, size, size), num_channels=64, num_res_blocks=1, precision="bf16", **kw)
m.load_state_dict(params); m = m.cuda().eval()
x = torch.randn(5, 3, size, size); t = torch.tensor(0.61)
want = O.wrapper_forward(cfg, params, t, x)
got = m(t.cuda(), x.cuda()).cpu()
print(sys.argv[1], CFGS[int(sys.argv[1])], "rel-L2", float((got - want).norm() / want.norm()), flush=True)
