"""NFE timing + per-kernel-class breakdown for every BASELINE.json config (bf16 engine), one GPU."""
import sys, time, collections, json, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as g
g.build(); pkg = g.load_package()
from oracle import unet as O

CONFIGS = {
    "0 mnist_cfm_inpaint b64": (O.config_from_wrapper((1, 28, 28), 32, 1, extra_in_channels=1), 64, 1),
    "1 cifar_cfm b1024": (O.config_from_wrapper((3, 32, 32), 128, 2, channel_mult=[1, 2, 2, 2], num_heads=4, num_head_channels=64, attention_resolutions="16"), 1024, 0),
    "2 mnist_classcond b4096": (O.config_from_wrapper((1, 28, 28), 32, 1, class_cond=True, num_classes=10), 4096, 0),
    "3a ddpm_mnist_amortized b256": (O.config_from_create_model(image_size=28, in_channels=2, out_channels=1, num_channels=32, num_res_blocks=1, channel_mult="1, 2, 2", resblock_updown=True), 256, 1),
    "3b ddpm_flowers64 b128": (O.config_from_create_model(image_size=64, in_channels=3, out_channels=3, num_channels=128, num_res_blocks=1, resblock_updown=True, num_head_channels=64, use_scale_shift_norm=True, num_heads=4), 128, 0),
    "4 superres128 b32": (O.config_from_create_model(image_size=128, in_channels=6, out_channels=3, num_channels=128, num_res_blocks=1), 32, 3),
}
only = sys.argv[1:] 
out = {}
for name, (cfg, B, ncond) in CONFIGS.items():
    if only and not any(name.startswith(o) for o in only): continue
    params = O.seeded_params(cfg, 0)
    m = pkg.UNetModel(image_size=cfg.image_size, in_channels=cfg.in_channels, model_channels=cfg.model_channels, out_channels=cfg.out_channels,
                      num_res_blocks=cfg.num_res_blocks, attention_resolutions=cfg.attention_ds, channel_mult=cfg.channel_mult, num_classes=cfg.num_classes,
                      num_heads=cfg.num_heads, num_head_channels=cfg.num_head_channels, num_heads_upsample=cfg.num_heads_upsample,
                      use_scale_shift_norm=cfg.use_scale_shift_norm, resblock_updown=cfg.resblock_updown, precision="bf16")
    m.load_state_dict(params); m = m.cuda().eval()
    e = m.engine()
    cx = cfg.in_channels - ncond
    x = torch.randn(B, cx, cfg.image_size, cfg.image_size, device='cuda')
    cond = torch.randn(B, ncond, cfg.image_size, cfg.image_size, device='cuda') if ncond else None
    y = (torch.arange(B, device='cuda') % cfg.num_classes) if cfg.num_classes else None
    for _ in range(2): e.forward(x, 0.5, y=y, cond=cond)
    torch.cuda.synchronize(); t0 = time.time()
    n = 5
    for _ in range(n): e.forward(x, 0.5, y=y, cond=cond)
    torch.cuda.synchronize(); dt = (time.time() - t0) / n
    rows = e.profile_forward(x, 0.5, y=y, cond=cond, repeats=3)
    agg = collections.defaultdict(float)
    for r in rows: agg[r['kind']] += r['ms']
    tf = e.flops_per_sample * B / dt / 1e12
    print(f"{name:32s} NFE {dt*1e3:8.3f} ms  {B/dt:10.0f} sample-NFE/s  {tf:7.1f} TFLOP/s  tc_convs={e.tensor_core_convs:3d}  " + "  ".join(f"{k}={v:.3f}" for k, v in sorted(agg.items())), flush=True)
    out[name] = {"ms_per_nfe": dt * 1e3, "sample_nfe_per_s": B / dt, "tflops": tf, "tc_convs": e.tensor_core_convs, "by_kind_ms": dict(agg)}
    del m, e
    torch.cuda.empty_cache()
import os
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/config_sweep.json', 'w'), indent=1)
