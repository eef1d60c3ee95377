import sys, time, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as g
pkg = g.load_package()
from oracle import unet as O
cases = {
 "flowers64 b128": (O.config_from_create_model(image_size=64, in_channels=3, out_channels=3, num_channels=128, num_res_blocks=1, resblock_updown=True, num_head_channels=64, use_scale_shift_norm=True, num_heads=4), 128, 0),
 "superres128 b32": (O.config_from_create_model(image_size=128, in_channels=6, out_channels=3, num_channels=128, num_res_blocks=1), 32, 3),
 "mnist_ddpm b256": (O.config_from_create_model(image_size=28, in_channels=2, out_channels=1, num_channels=32, num_res_blocks=1, channel_mult="1, 2, 2", resblock_updown=True), 256, 1),
}
for name, (cfg, B, ncond) in cases.items():
    for fuse in (True, False):
        m = pkg.UNetModel(image_size=cfg.image_size, in_channels=cfg.in_channels, model_channels=cfg.model_channels, out_channels=cfg.out_channels,
                          num_res_blocks=cfg.num_res_blocks, attention_resolutions=cfg.attention_ds, channel_mult=cfg.channel_mult, num_classes=cfg.num_classes,
                          num_heads=cfg.num_heads, num_head_channels=cfg.num_head_channels, num_heads_upsample=cfg.num_heads_upsample,
                          use_scale_shift_norm=cfg.use_scale_shift_norm, resblock_updown=cfg.resblock_updown, precision="bf16", fuse_groupnorm=fuse)
        m.load_state_dict(O.seeded_params(cfg, 0)); m = m.cuda().eval()
        e = m.engine()
        x = torch.randn(B, cfg.in_channels - ncond, cfg.image_size, cfg.image_size, device='cuda')
        cond = torch.randn(B, ncond, cfg.image_size, cfg.image_size, device='cuda') if ncond else None
        for _ in range(3): e.forward(x, 0.5, cond=cond)
        torch.cuda.synchronize(); t0 = time.time()
        for _ in range(10): e.forward(x, 0.5, cond=cond)
        torch.cuda.synchronize(); dt = (time.time() - t0) / 10
        rows = e.profile_forward(x, 0.5, cond=cond, repeats=3)
        nf = sum('+' in r['name'] for r in rows)
        print(f"{name:18s} fuse={fuse!s:5s} NFE {dt*1e3:7.3f} ms  launches={e.last_launches} folded={nf}")
        if fuse:
            for r in rows:
                if '+' in r['name']: print(f"      {r['name']:70s} {r['ms']:.4f}")
        del m, e
