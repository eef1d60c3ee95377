"""Debug helper: one NFE of a golden config in bf16 against the committed golden output, under env toggles."""
import os, sys
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as g
pkg = g.load_package()
from oracle import unet as O
from golden_configs import GOLDEN_CONFIGS
from test_gpu_unet import build, rel_l2
for name in sys.argv[1:] or ["cifar"]:
    cfg, _, _ = GOLDEN_CONFIGS[name]
    gd = np.load(f"tests/golden/unet_{name}.npz")
    params = O.seeded_params(cfg, int(gd["seed"]))
    m = build(pkg, cfg, params, "bf16", "cuda")
    out = m(torch.from_numpy(gd["x"]).cuda(), torch.from_numpy(gd["t"]).cuda()).cpu()
    print(name, "rel-L2", rel_l2(out, torch.from_numpy(gd["out"])), "env", {k: v for k, v in os.environ.items() if k.startswith("CFM_")})
