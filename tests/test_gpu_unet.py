"""GPU parity of one NFE: the CUDA engine (through the C ABI) against the CPU oracle and the
committed golden vectors of the reference's own code.

Tolerances (north_star): fp32 mode <= 1e-4 relative L2 per NFE; bf16 mode <= 2e-2."""
import os

import numpy as np
import pytest
import torch

from golden_configs import GOLDEN_CONFIGS, random_config
from oracle import unet as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = {"fp32": 1e-4, "bf16": 2e-2}


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def build(pkg, cfg, params, precision, dev, **extra):
    m = pkg.UNetModel(image_size=cfg.image_size, in_channels=cfg.in_channels, model_channels=cfg.model_channels,
                      out_channels=cfg.out_channels, num_res_blocks=cfg.num_res_blocks,
                      attention_resolutions=cfg.attention_ds, channel_mult=cfg.channel_mult, num_classes=cfg.num_classes,
                      num_heads=cfg.num_heads, num_head_channels=cfg.num_head_channels,
                      num_heads_upsample=cfg.num_heads_upsample, use_scale_shift_norm=cfg.use_scale_shift_norm,
                      resblock_updown=cfg.resblock_updown, use_new_attention_order=cfg.use_new_attention_order,
                      precision=precision, **extra)
    m.load_state_dict(params)
    return m.to(dev).eval()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", list(GOLDEN_CONFIGS))
def test_nfe_matches_reference_golden(pkg, cuda, name, precision):
    cfg, _, _ = GOLDEN_CONFIGS[name]
    g = np.load(os.path.join(GOLD, f"unet_{name}.npz"))
    params = O.seeded_params(cfg, int(g["seed"]))
    m = build(pkg, cfg, params, precision, cuda)
    out = m(torch.from_numpy(g["x"]).to(cuda), torch.from_numpy(g["t"]).to(cuda)).cpu()
    assert m.engine().param_count == int(g["n_params"])
    r = rel_l2(out, torch.from_numpy(g["out"]))
    print(f"{name}[{precision}] rel-L2 = {r:.3e}")
    assert r < TOL[precision], r


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_wrapper_scalar_t_and_class_labels(pkg, cuda, precision):
    # conditional_mnist.ipynb cell 2-4: UNetModel(dim=(1,28,28), num_channels=32, num_res_blocks=1, num_classes=10, class_cond=True)
    cfg = O.config_from_wrapper((1, 28, 28), 32, 1, class_cond=True, num_classes=10)
    params = O.seeded_params(cfg, 21)
    m = pkg.UNetModelWrapper(dim=(1, 28, 28), num_channels=32, num_res_blocks=1, num_classes=10, class_cond=True,
                             precision=precision)
    m.load_state_dict(params)
    m = m.to(cuda).eval()
    x = torch.randn(20, 1, 28, 28)
    y = torch.arange(10).repeat(2)
    t0 = torch.tensor(0.4321)
    want = O.wrapper_forward(cfg, params, t0, x, y)
    got = m(t0.to(cuda), x.to(cuda), y.to(cuda)).cpu()           # 0-dim t: shared-row embedding path
    assert rel_l2(got, want) < TOL[precision]
    tb = torch.rand(20)
    want = O.wrapper_forward(cfg, params, tb, x, y)
    got = m(tb.to(cuda), x.to(cuda), y.to(cuda), args={}).cpu()  # per-sample t; torchdyn's extra kwarg tolerated
    assert rel_l2(got, want) < TOL[precision]
    with pytest.raises(AssertionError):
        m(t0.to(cuda), x.to(cuda))                               # y required iff class conditional


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_inpaint_and_superres_wrappers(pkg, cuda, precision):
    cfg = O.config_from_wrapper((1, 28, 28), 32, 1, extra_in_channels=1)
    params = O.seeded_params(cfg, 22)
    m = pkg.InPaintModelWrapper(dim=(1, 28, 28), num_channels=32, num_res_blocks=1, num_classes=None, class_cond=True,
                                precision=precision)
    m.load_state_dict(params)
    m = m.to(cuda).eval()
    x, con = torch.randn(3, 1, 28, 28), torch.rand(3, 1, 28, 28) * 2 - 1
    con[:, :, 6:20, 7:21] = -2.0
    t = torch.tensor(0.25)
    want = O.inpaint_forward(cfg, params, x, t, con)
    got = m.forward(x.to(cuda), t.to(cuda), con=con.to(cuda)).cpu()     # (x, t) order: utils_mnist.py:97
    assert rel_l2(got, want) < TOL[precision]

    cfg = O.config_from_wrapper((3, 32, 32), 32, 1, extra_in_channels=3, channel_mult=(1, 2))
    params = O.seeded_params(cfg, 23)
    s = pkg.SuperResModelWrapper(dim=(3, 32, 32), num_channels=32, num_res_blocks=1, channel_mult=(1, 2),
                                 num_classes=None, class_cond=True, precision=precision)
    s.load_state_dict(params)
    s = s.to(cuda).eval()
    x, lo = torch.randn(2, 3, 32, 32), torch.rand(2, 3, 8, 8)
    want = O.superres_forward(cfg, params, x, t, lo)
    got = s.forward(x.to(cuda), t.to(cuda), low_res=lo.to(cuda)).cpu()
    assert rel_l2(got, want) < TOL[precision]


def test_batch_independence_and_ragged_batches(pkg, cuda):
    # samples are independent (per-sample GroupNorm / attention): any batch split gives the same rows
    cfg, _, _ = GOLDEN_CONFIGS["tiny_neworder"]
    params = O.seeded_params(cfg, 5)
    for precision in ("fp32", "bf16"):
        m = build(pkg, cfg, params, precision, cuda)
        x = torch.randn(37, 3, 16, 16, device=cuda)
        t = torch.rand(37, device=cuda)
        full = m(x, t)
        parts = torch.cat([m(x[:1], t[:1]), m(x[1:20], t[1:20]), m(x[20:], t[20:])])
        assert torch.equal(full, parts)
    with pytest.raises(ValueError):
        m(torch.randn(2, 3, 15, 16, device=cuda), torch.rand(2, device=cuda))
    # empty batch (a rank with no samples): an empty result like the PyTorch module, from the forward and the samplers
    empty = m(torch.zeros(0, 3, 16, 16, device=cuda), torch.zeros(0, device=cuda))
    assert tuple(empty.shape) == (0, 3, 16, 16)
    e = m.engine()
    xe, traj, img = e.sample_euler(torch.zeros(0, 3, 16, 16, device=cuda), [0.0, 0.5], [0.5, 0.5], return_trajectory=True,
                                   return_uint8=True)
    assert tuple(xe.shape) == (0, 3, 16, 16) and tuple(traj.shape) == (3, 0, 3, 16, 16) and img.numel() == 0


def test_engine_flops_and_param_accounting(pkg, cuda):
    cfg, _, _ = GOLDEN_CONFIGS["cifar"]
    m = build(pkg, cfg, O.seeded_params(cfg, 0), "bf16", cuda)
    e = m.engine()
    assert e.param_count == 35_746_307
    assert abs(e.flops_per_sample / 1e9 - 12.444) < 0.01
    m(torch.randn(1, 3, 32, 32, device=cuda), torch.rand(1, device=cuda))
    assert e.last_launches > 0


def _kinds(m, x, t):
    return {r["kind"] for r in m.engine().profile_forward(x, float(t), repeats=1)}


@pytest.mark.parametrize("name", ["mnist_cfm", "mnist_ddpm", "flowers_ddpm", "cifar"])
def test_bf16_runs_on_tensor_core_kernels(pkg, cuda, name):
    # every conv / attention of the BASELINE configs must take the tcgen05 kernels in bf16 mode - a silent fall
    # back to the CUDA-core kernels would keep parity green and lose two orders of magnitude
    cfg, _, _ = GOLDEN_CONFIGS[name]
    m = build(pkg, cfg, O.seeded_params(cfg, 1), "bf16", cuda)
    x = torch.randn(2, cfg.in_channels, cfg.image_size, cfg.image_size, device=cuda)
    kinds = _kinds(m, x, 0.3)
    assert "conv_generic" not in kinds and "attention_generic" not in kinds, kinds
    assert "conv_tcgen05" in kinds and "attention" in kinds


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("size,mult,heads,hc,attn", [(20, (1, 2), -1, 32, "1,2"), (12, (1, 1, 2), 2, -1, "1"), (24, (2, 1), -1, 64, "2"),
                                                      (16, (2, 4, 6), 1, -1, "1,2"), (8, (3, 6), -1, 192, "1"), (8, (8,), 1, -1, "1")])
def test_odd_maps_and_head_widths_match_oracle(pkg, cuda, precision, size, mult, heads, hc, attn):
    # maps that are not powers of two (20/10, 12/6/3, 24/12), sequence lengths 400/100/144, 32- and 64-wide heads,
    # strided and folded-upsample convs on them: the flash attention and the generalised conv tiles against the oracle;
    # the last two: single 128 / 256 / 384-wide heads over 256 / 64 / 16 tokens and 192 / 512-wide heads (wide-head kernel)
    kw = dict(channel_mult=list(mult), attention_resolutions=",".join(str(size // int(a)) for a in attn.split(",")))
    if hc > 0: kw.update(num_head_channels=hc)
    else: kw.update(num_heads=heads)
    cfg = O.config_from_wrapper((3, size, size), 64, 1, **kw)
    params = O.seeded_params(cfg, 31)
    m = pkg.UNetModelWrapper(dim=(3, size, size), num_channels=64, num_res_blocks=1, precision=precision, **kw)
    m.load_state_dict(params)
    m = m.to(cuda).eval()
    x = torch.randn(5, 3, size, size)
    t = torch.tensor(0.61)
    want = O.wrapper_forward(cfg, params, t, x)
    got = m(t.to(cuda), x.to(cuda)).cpu()
    r = rel_l2(got, want)
    print(f"odd[{size},{mult},{precision}] rel-L2 = {r:.3e}")
    assert r < TOL[precision], r


@pytest.mark.parametrize("seed", range(16))
def test_random_configs_match_oracle(pkg, cuda, seed):
    # seeded sweep over the constructor's keyword space: channel multipliers 1-3 on 32 / 64 / 96 base channels (K-iterations
    # of 32 and 64 channels, N tiles of every width), 1-3 levels, FiLM, up/down ResBlocks, both attention orders, head
    # counts / widths, class labels, ragged batches - both precisions against the CPU oracle
    kw, B = random_config(seed)
    cfg = O.config_from_create_model(**kw)
    params = O.seeded_params(cfg, 900 + seed)
    rs = np.random.RandomState(seed)
    x = torch.from_numpy(rs.standard_normal((B, kw["in_channels"], kw["image_size"], kw["image_size"])).astype(np.float32))
    t = torch.from_numpy(rs.uniform(0, 1, size=(B,)).astype(np.float32))
    y = torch.from_numpy(rs.randint(0, kw["num_classes"], size=(B,))) if "num_classes" in kw else None
    want = O.unet_forward(cfg, params, x, t, y)
    for precision in ("fp32", "bf16"):
        m = build(pkg, cfg, params, precision, cuda)
        got = (m(x.to(cuda), t.to(cuda)) if y is None else m(x.to(cuda), t.to(cuda), y.to(cuda))).cpu()
        r = rel_l2(got, want)
        print(f"random[{seed}] {kw} B={B} [{precision}] rel-L2 = {r:.3e}")
        assert r < TOL[precision], (kw, B, precision, r)


def test_superres_config_runs_on_tensor_core_kernels(pkg, cuda):
    # BASELINE config 5 (4x super-resolution, 128x128, low-res image concatenated): single 384 / 512-wide heads take the
    # wide-head attention kernel, everything else the same tcgen05 kernels; checked against the oracle at batch 1
    cfg = O.config_from_create_model(image_size=128, in_channels=6, out_channels=3, num_channels=128, num_res_blocks=1)
    params = O.seeded_params(cfg, 3)
    m = build(pkg, cfg, params, "bf16", cuda)
    x = torch.randn(1, 6, 128, 128)
    t = torch.tensor([0.4])
    kinds = _kinds(m, x.to(cuda), 0.4)
    assert "conv_generic" not in kinds and "attention_generic" not in kinds, kinds
    want = O.unet_forward(cfg, params, x, t)
    got = m(x.to(cuda), t.to(cuda)).cpu()
    r = rel_l2(got, want)
    print(f"superres128[bf16] rel-L2 = {r:.3e}")
    assert r < TOL["bf16"], r


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_engines_on_two_devices_in_one_process(pkg):
    # kernel attributes (dynamic shared memory) are per device context: a second engine on another GPU of the same
    # process must set them again; both devices must give the same bits
    outs = []
    for name in ("cifar", "mnist_ddpm"):
        cfg, _, _ = GOLDEN_CONFIGS[name]
        g = np.load(os.path.join(GOLD, f"unet_{name}.npz"))
        params = O.seeded_params(cfg, int(g["seed"]))
        for precision in ("bf16", "fp32"):
            per_dev = []
            for dev in ("cuda:0", "cuda:1"):
                m = build(pkg, cfg, params, precision, torch.device(dev))
                per_dev.append(m(torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["t"]).to(dev)).cpu())
            assert torch.equal(per_dev[0], per_dev[1]), (name, precision)
            assert rel_l2(per_dev[1], torch.from_numpy(g["out"])) < TOL[precision]


def test_engines_on_two_host_threads(pkg, cuda):
    # the tensor-map caches of the attention kernels are process-wide; two engines driven from two host
    # threads (own CUDA streams) must not corrupt them and must give the bits a single thread gives
    import threading
    jobs = []
    for name in ("cifar", "tiny_neworder", "mnist_ddpm"):
        cfg, _, _ = GOLDEN_CONFIGS[name]
        g = np.load(os.path.join(GOLD, f"unet_{name}.npz"))
        params = O.seeded_params(cfg, int(g["seed"]))
        x, t = torch.from_numpy(g["x"]).to(cuda), torch.from_numpy(g["t"]).to(cuda)
        want = build(pkg, cfg, params, "bf16", cuda)(x, t).cpu()
        jobs.append((cfg, params, x, t, want))
    results, errors = {}, []

    def work(i):
        try:
            cfg, params, x, t, _ = jobs[i % len(jobs)]
            with torch.cuda.stream(torch.cuda.Stream(device=cuda)):
                m = build(pkg, cfg, params, "bf16", cuda)        # a fresh engine per thread: first launches fill the caches
                outs = [m(x.repeat(b, 1, 1, 1), t.repeat(b))[:x.shape[0]].cpu() for b in (1, 2, 3)]
            results[i] = outs
        except Exception as ex:                                   # noqa: BLE001
            errors.append(repr(ex))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(6)]
    for th in threads: th.start()
    for th in threads: th.join()
    assert not errors, errors
    for i, outs in results.items():
        for o in outs:
            assert torch.equal(o, jobs[i % len(jobs)][4]), i


@pytest.mark.parametrize("B", [128, 1000, 1024, 2048])
def test_cifar_nfe_at_benchmarked_batch(pkg, cuda, B):
    # the shape bench.py times (CIFAR bf16, 1024 samples per NFE): multi-wave persistent tiling, stationary-weight pair
    # counts, cluster GroupNorm and arena offsets beyond 2^31 bytes are only reached here.  The golden rows are planted
    # at the start, in the middle and at the end of the batch (1000 is not a multiple of any tile count); every copy
    # must equal the B = 2 result bit for bit and sit within the bf16 bar of the reference's own output.  2048: a 4.8 GiB
    # arena - byte offsets beyond 2^32.
    cfg, _, _ = GOLDEN_CONFIGS["cifar"]
    g = np.load(os.path.join(GOLD, "unet_cifar.npz"))
    params = O.seeded_params(cfg, int(g["seed"]))
    m = build(pkg, cfg, params, "bf16", cuda)
    gx, gt = torch.from_numpy(g["x"]).to(cuda), torch.from_numpy(g["t"]).to(cuda)
    small = m(gx, gt)
    gen = torch.Generator(device=cuda).manual_seed(B)
    x = torch.randn(B, 3, 32, 32, device=cuda, generator=gen)
    t = torch.rand(B, device=cuda, generator=gen)
    spots = [0, B // 2 - 1, B - 2]
    for s in spots:
        x[s:s + 2] = gx; t[s:s + 2] = gt
    out = m(x, t)
    assert torch.isfinite(out).all()
    for s in spots:
        assert torch.equal(out[s:s + 2], small), f"rows {s}..{s + 1} of a batch of {B} differ from the batch-2 result"
    r = rel_l2(out[:2].cpu(), torch.from_numpy(g["out"]))
    print(f"cifar[bf16, B={B}] rel-L2 vs reference golden = {r:.3e}")
    assert r < TOL["bf16"], r
    # uniform (scalar) t as the sampler uses it: the shared-row embedding path at the same batch
    u = m(x, gt[0])
    assert torch.equal(u[:1], m(gx[:1], gt[0]))
    assert m.engine().workspace_bytes(B) == m.engine().workspace_bytes(1) * B


@pytest.mark.parametrize("name,B", [("mnist_inpaint", 64), ("mnist_cfm", 4096), ("mnist_ddpm", 256), ("flowers_ddpm", 128)])
def test_other_configs_at_their_benchmarked_batch(pkg, cuda, name, B):
    # BASELINE configs 0, 2, 3a, 3b at the batch their throughput is quoted on (profiles/sampler_sweep.py): the golden
    # rows of the reference's own evaluation are planted at the start, the middle and the end of a random batch; every
    # copy must equal the small-batch result bit for bit (size-independent property: a sample's result does not depend on
    # the batch it sits in) and stay within the bf16 bar of the reference output.
    cfg, _, _ = GOLDEN_CONFIGS[name]
    g = np.load(os.path.join(GOLD, f"unet_{name}.npz"))
    params = O.seeded_params(cfg, int(g["seed"]))
    m = build(pkg, cfg, params, "bf16", cuda)
    gx, gt = torch.from_numpy(g["x"]).to(cuda), torch.from_numpy(g["t"]).to(cuda)
    n = gx.shape[0]
    small = m(gx, gt)
    gen = torch.Generator(device=cuda).manual_seed(B)
    x = torch.randn(B, *gx.shape[1:], device=cuda, generator=gen)
    t = torch.rand(B, device=cuda, generator=gen)
    spots = [0, B // 2 - 1, B - n]
    for s in spots:
        x[s:s + n] = gx; t[s:s + n] = gt
    out = m(x, t)
    assert torch.isfinite(out).all()
    for s in spots:
        assert torch.equal(out[s:s + n], small), f"{name}: rows {s}..{s + n - 1} of a batch of {B} differ from the batch-{n} result"
    r = rel_l2(out[:n].cpu(), torch.from_numpy(g["out"]))
    print(f"{name}[bf16, B={B}] rel-L2 vs reference golden = {r:.3e}")
    assert r < TOL["bf16"], r


@pytest.mark.parametrize("name,batch", [("cifar", 5), ("flowers_ddpm", 2), ("tiny_neworder", 7)])
def test_groupnorm_folded_into_conv_epilogue(pkg, cuda, name, batch):
    # default path: a ResBlock's out_layers.0/1 (GroupNorm + SiLU) runs in the epilogue of its first conv (conv_tc2_kernel
    # <.., true>: statistics from the fp32 accumulators, samples that span several CTA tiles exchange partial sums through
    # L2); fuse_groupnorm=False (CFM_FLAG_SEPARATE_GROUPNORM) keeps the separate GroupNorm pass.  Both must sit within the bf16 bar of the reference
    # golden and close to each other.  cifar covers 32x32 (4 tiles per sample), 16x16 (2), 8x8 and 4x4 maps (2 / 8 samples
    # per tile); flowers uses FiLM and must NOT fold; batch position must not matter (ragged tail tiles included).
    cfg, _, _ = GOLDEN_CONFIGS[name]
    g = np.load(os.path.join(GOLD, f"unet_{name}.npz"))
    params = O.seeded_params(cfg, int(g["seed"]))
    gx, gt = torch.from_numpy(g["x"]), torch.from_numpy(g["t"])
    reps = (batch + gx.shape[0] - 1) // gx.shape[0]
    x = torch.cat([gx] * reps)[:batch].to(cuda)
    t = torch.cat([gt] * reps)[:batch].to(cuda)
    m = build(pkg, cfg, params, "bf16", cuda)
    fused = m(x, t).cpu()
    names = [r["name"] for r in m.engine().profile_forward(x, 0.5, repeats=1)]
    n_folded = sum("+" in n for n in names)                 # conv1 + out_layers.0, and block outputs + the next GroupNorm
    assert torch.equal(m(x, t).cpu(), fused)             # a second evaluation (next epoch of the exchange flags): same bits
    m2 = build(pkg, cfg, params, "bf16", cuda, fuse_groupnorm=False)
    plain = m2(x, t).cpu()
    names2 = [r["name"] for r in m2.engine().profile_forward(x, 0.5, repeats=1)]
    assert not any("+" in n for n in names2) and len(names2) == len(names) + n_folded
    if name == "cifar":
        assert sum("+out_layers.0" in n for n in names) == 22, names
        assert n_folded >= 22 + 10, names                    # attention norms, in_layers.0 of single-source blocks, out.0
    if name == "flowers_ddpm":
        assert sum("+out_layers.0" in n for n in names) == 0    # FiLM blocks keep out_layers.0 as a pass of its own
    want = torch.from_numpy(g["out"])
    n = want.shape[0]
    r_f, r_p, d = rel_l2(fused[:n], want), rel_l2(plain[:n], want), rel_l2(fused, plain)
    print(f"{name}: {n_folded} GroupNorms folded; rel-L2 folded {r_f:.3e}, separate pass {r_p:.3e}, between them {d:.3e}")
    assert r_f < TOL["bf16"] and r_p < TOL["bf16"] and d < TOL["bf16"]
    for k in range(n, batch):                       # repeated golden rows: batch position must not matter
        assert torch.equal(fused[k], fused[k % n])
