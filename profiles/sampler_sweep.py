"""End-to-end sampling throughput of every BASELINE.json config through the public Python API (bf16, CUDA graphs,
synthetic data, seeded weights).  One GPU, or N GPUs under torchrun (one rank per GPU, the per-GPU batch of the config
on every rank = weak scaling, no collective in the loop; time = CUDA events, max over ranks; samples/s = whole job).
Prints one line per config and writes gpurun_out/sampler_sweep[_nN].json.
usage: python profiles/sampler_sweep.py [config-prefix ...]
       python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/sampler_sweep.py"""
import json, os, sys, time
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as g
RANK, WORLD, LOCAL = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(LOCAL)
if WORLD > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
if RANK == 0: g.build()
if WORLD > 1: dist.barrier()
pkg = g.load_package()
from oracle import unet as O


def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()          # warm-up (graph capture, tensor maps)
    if WORLD > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device='cuda')
    if WORLD > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms) / 1e3


def wrapper(cls, cfg_kw, oracle_cfg, precision="bf16", **kw):
    m = cls(precision=precision, **cfg_kw, **kw)
    m.load_state_dict(O.seeded_params(oracle_cfg, 0))
    return m.cuda().eval()


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def drift(run_bf16, run_fp32):
    """Final-sample drift of the bf16 sampler against the fp32 (exact-mode) engine on the same inputs / noise seed."""
    if RANK: return float('nan')
    a, b = run_bf16(), run_fp32()
    return rel(a, b)


out = {}
only = sys.argv[1:]
def want(name): return not only or any(name.startswith(o) for o in only)

if want("0"):
    B = 64
    cfg = O.config_from_wrapper((1, 28, 28), 32, 1, extra_in_channels=1)
    m = wrapper(pkg.InPaintModelWrapper, dict(dim=(1, 28, 28), num_channels=32, num_res_blocks=1), cfg, num_classes=None, class_cond=True)
    x0 = torch.randn(B, 1, 28, 28, device='cuda'); con = torch.rand(B, 1, 28, 28, device='cuda') * 2 - 1; con[:, :, 6:20, 7:21] = -2
    ts = torch.linspace(0, 1, 101)
    dt = timed(lambda: pkg.sample_euler(m, x0, ts, cond=con, cond_drift=True, use_graph=True))
    m32 = wrapper(pkg.InPaintModelWrapper, dict(dim=(1, 28, 28), num_channels=32, num_res_blocks=1), cfg, precision="fp32", num_classes=None, class_cond=True)
    d = drift(lambda: pkg.sample_euler(m, x0[:4], ts, cond=con[:4], cond_drift=True), lambda: pkg.sample_euler(m32, x0[:4], ts, cond=con[:4], cond_drift=True))
    out["0 mnist_cfm_inpaint b64 100-step Euler"] = {"s": dt, "samples_per_s": WORLD * B / dt, "n_gpus": WORLD, "evals": 100, "drift_bf16_vs_fp32": d}
if want("2"):
    B = 4096
    cfg = O.config_from_wrapper((1, 28, 28), 32, 1, class_cond=True, num_classes=10)
    m = wrapper(pkg.UNetModelWrapper, dict(dim=(1, 28, 28), num_channels=32, num_res_blocks=1), cfg, num_classes=10, class_cond=True)
    x0 = torch.randn(B, 1, 28, 28, device='cuda'); y = torch.arange(B, device='cuda') % 10
    ts = torch.linspace(0, 1, 101)
    dt = timed(lambda: pkg.sample_euler(m, x0, ts, y=y, guidance_weight=2.0, use_graph=True), reps=1)
    m32 = wrapper(pkg.UNetModelWrapper, dict(dim=(1, 28, 28), num_channels=32, num_res_blocks=1), cfg, precision="fp32", num_classes=10, class_cond=True)
    d = drift(lambda: pkg.sample_euler(m, x0[:8], ts, y=y[:8], guidance_weight=2.0), lambda: pkg.sample_euler(m32, x0[:8], ts, y=y[:8], guidance_weight=2.0))
    out["2 mnist_classcond CFG b4096 100-step Euler x2 evals"] = {"s": dt, "samples_per_s": WORLD * B / dt, "n_gpus": WORLD, "evals": 200, "drift_bf16_vs_fp32": d}
if want("3a"):
    B = 256
    cfg = O.config_from_create_model(image_size=28, in_channels=2, out_channels=1, num_channels=32, num_res_blocks=1, channel_mult="1, 2, 2", resblock_updown=True)
    net = pkg.create_model(image_size=28, in_channels=2, out_channels=1, num_channels=32, num_res_blocks=1, channel_mult="1, 2, 2", resblock_updown=True, precision="bf16")
    net.load_state_dict(O.seeded_params(cfg, 0)); net = net.cuda().eval()
    ddpm = pkg.DDPM(1000)
    lik = pkg.InPainting(14, -2.0)
    fn = pkg.get_conditional_sample_fn(lambda xi, i: net(xi, 1.0 * i / ddpm.Ns), ddpm, pkg.Amortized(), lik, use_graph=True)   # main.py:140 verbatim
    xT = torch.randn(B, 1, 28, 28, device='cuda'); cond = lik.sample(torch.rand(B, 1, 28, 28, device="cuda") * 2 - 1)
    dt = timed(lambda: fn(xT, cond), reps=1)
    net32 = pkg.create_model(image_size=28, in_channels=2, out_channels=1, num_channels=32, num_res_blocks=1, channel_mult="1, 2, 2", resblock_updown=True, precision="fp32")
    net32.load_state_dict(O.seeded_params(cfg, 0)); net32 = net32.cuda().eval()
    f16 = pkg.get_conditional_sample_fn(pkg.EpsModel(net, ddpm), ddpm, pkg.Amortized(), lik, seed=5)
    f32 = pkg.get_conditional_sample_fn(pkg.EpsModel(net32, ddpm), ddpm, pkg.Amortized(), lik, seed=5)
    d = drift(lambda: f16(xT[:4], cond[:4]), lambda: f32(xT[:4], cond[:4]))
    out["3a ddpm_mnist amortized inpainting b256 1000 steps"] = {"s": dt, "samples_per_s": WORLD * B / dt, "n_gpus": WORLD, "evals": 1000, "drift_bf16_vs_fp32": d}
if want("3b"):
    B = 128
    kw = dict(image_size=64, in_channels=3, out_channels=3, num_channels=128, num_res_blocks=1, resblock_updown=True, num_head_channels=64, use_scale_shift_norm=True, num_heads=4)
    cfg = O.config_from_create_model(**kw)
    net = pkg.create_model(precision="bf16", **kw)
    net.load_state_dict(O.seeded_params(cfg, 0)); net = net.cuda().eval()
    ddpm = pkg.DDPM(1000)
    lik = pkg.InPainting(20, -2.0)
    fn = pkg.get_conditional_sample_fn(lambda xi, i: net(xi, 1.0 * i / ddpm.Ns), ddpm, pkg.Replacement(), lik, use_graph=True)  # main.py:140 verbatim
    xT = torch.randn(B, 3, 64, 64, device='cuda'); cond = lik.sample(torch.rand(B, 3, 64, 64, device="cuda") * 2 - 1)
    dt = timed(lambda: fn(xT, cond), reps=1)
    net32 = pkg.create_model(precision="fp32", **kw)
    net32.load_state_dict(O.seeded_params(cfg, 0)); net32 = net32.cuda().eval()
    d40 = pkg.DDPM(40)                       # drift on a 40-step chain, 2 samples (the exact-mode engine runs on CUDA cores)
    f16 = pkg.get_conditional_sample_fn(pkg.EpsModel(net, d40), d40, pkg.Replacement(), lik, seed=5)
    f32 = pkg.get_conditional_sample_fn(pkg.EpsModel(net32, d40), d40, pkg.Replacement(), lik, seed=5)
    d = drift(lambda: f16(xT[:2], cond[:2]), lambda: f32(xT[:2], cond[:2]))
    out["3b ddpm_flowers64 RePaint-style replacement b128 1000 steps"] = {"s": dt, "samples_per_s": WORLD * B / dt, "n_gpus": WORLD, "evals": 1000, "drift_bf16_vs_fp32": d, "drift_chain_steps": 40}
if want("4"):
    B = 32
    cfg = O.config_from_wrapper((3, 128, 128), 128, 1, extra_in_channels=3)
    m = wrapper(pkg.SuperResModelWrapper, dict(dim=(3, 128, 128), num_channels=128, num_res_blocks=1), cfg, num_classes=None, class_cond=True)
    x0 = torch.randn(B, 3, 128, 128, device='cuda'); lo = torch.rand(B, 3, 32, 32, device='cuda') * 2 - 1
    up = pkg.resize_bilinear(lo, (128, 128))
    ts = torch.linspace(0, 1, 51)
    dt = timed(lambda: pkg.sample_euler(m, x0, ts, cond=up, use_graph=True))
    m32 = wrapper(pkg.SuperResModelWrapper, dict(dim=(3, 128, 128), num_channels=128, num_res_blocks=1), cfg, precision="fp32", num_classes=None, class_cond=True)
    t10 = torch.linspace(0, 1, 11)
    d = drift(lambda: pkg.sample_euler(m, x0[:1], t10, cond=up[:1]), lambda: pkg.sample_euler(m32, x0[:1], t10, cond=up[:1]))
    out["4 superres 32->128 CFM b32 50-step Euler"] = {"s": dt, "samples_per_s": WORLD * B / dt, "n_gpus": WORLD, "evals": 50, "drift_bf16_vs_fp32": d, "drift_steps": 10}

for k, v in out.items():
    if RANK: break
    print(f"[{WORLD} GPU] " + f"{k:62s} {v['s']:8.3f} s  {v['samples_per_s']:10.1f} samples/s  {v['s'] / v['evals'] * 1e3:8.3f} ms per U-Net eval batch  "
          f"final-sample drift bf16 vs fp32 engine {v.get('drift_bf16_vs_fp32', float('nan')):.2e}", flush=True)
if RANK == 0:
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(out, open('gpurun_out/sampler_sweep' + (f'_n{WORLD}' if WORLD > 1 else '') + '.json', 'w'), indent=1)
if WORLD > 1:
    dist.barrier(); dist.destroy_process_group()
