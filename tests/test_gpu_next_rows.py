"""GPU parity of the rows either side of the hot loop: plain-callable eps networks, per-call noise, the
Euler-Maruyama samplers, bilinear condition construction, FID statistics, and the hygiene items around them
(bounded graph cache, deterministic error norm)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from golden_configs import CHAIN_NS, chain_inputs, chain_net_cfg
from oracle import ddpm as D
from oracle import integrators as I
from oracle import unet as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def chain_net(pkg, cuda, in_ch, precision="fp32"):
    cfg = chain_net_cfg(in_ch)
    params = O.seeded_params(cfg, 41)
    m = pkg.create_model(image_size=16, in_channels=in_ch, out_channels=1, num_channels=32, num_res_blocks=1,
                         channel_mult="1,2", attention_resolutions="8", resblock_updown=True, precision=precision)
    m.load_state_dict(params)
    return cfg, params, m.to(cuda).eval()


def test_main_py_lambda_runs_unmodified(pkg, cuda):
    """AD/experiments/main.py:140-142 verbatim: `eps_model_ema = lambda xi, i: ema_network(xi, 1.0 * i / ddpm.Ns)` handed to
    get_conditional_sample_fn.  The closure is recognised as the engine's own forward (two-point probe) and the chain runs
    natively: same bits as the EpsModel form; and the oracle chain on the same injected noise within fp32 tolerance."""
    cfg, params, ema_network = chain_net(pkg, cuda, 1)
    ddpm = pkg.DDPM(CHAIN_NS)
    xT, cond = chain_inputs(103)
    noise = torch.randn(CHAIN_NS, 2, xT.numel(), generator=torch.Generator().manual_seed(5))
    conditioning, likelihood = pkg.Replacement(start_fraction=0.75), pkg.InPainting(6, -2.0)
    eps_model_ema = lambda xi, i: ema_network(xi, 1.0 * i / ddpm.Ns)
    cond_sample_fn = pkg.get_conditional_sample_fn(eps_model_ema, ddpm, conditioning, likelihood, noise=noise.to(cuda))
    got = cond_sample_fn(xT.to(cuda), cond.to(cuda))
    ref_fn = pkg.get_conditional_sample_fn(pkg.EpsModel(ema_network, ddpm), ddpm, conditioning, likelihood, noise=noise.to(cuda))
    assert torch.equal(got, ref_fn(xT.to(cuda), cond.to(cuda)))
    eps_cpu = lambda xi, t: O.unet_forward(cfg, params, xi, t)
    want = _oracle_replacement_with_slots(eps_cpu, CHAIN_NS, xT, cond, noise.view(CHAIN_NS, 2, *xT.shape), 0.75)
    assert rel_l2(got.cpu(), want) < 2e-4


def _oracle_replacement_with_slots(eps_model, Ns, xT, condition, slots, start_fraction, pad_value=-2.0):
    """oracle.ddpm.sample_replacement with the noise read from [Ns, 2, ...] slots (slot 0: q_sample draw of step i,
    slot 1: posterior draw) instead of a sequential stream - the layout cfm_sample_ddpm / cfm_ddpm_step document."""
    tb = D.ddpm_tables(Ns)
    xi = xT
    for i in reversed(range(Ns)):
        if i < int(Ns * start_fraction):
            nc = tb["sqrt_alphas_cumprod"][i] * condition + tb["sqrt_one_minus_alphas_cumprod"][i] * slots[i, 0]
            xi = torch.where(condition == pad_value, xi, nc)
        eps = eps_model(xi, D.eps_time(i, Ns).repeat(xi.shape[0]))
        xi, _ = D.posterior_step(tb, xi, eps, i, slots[i, 1] if i > 0 else 0.0)
    return torch.clip(xi, -1, 1)


class TorchEps(torch.nn.Module):
    """An eps network that is NOT engine-backed: plain PyTorch, so the sampler must take the stepwise path."""

    def __init__(self, in_ch):
        super().__init__()
        self.c1 = torch.nn.Conv2d(in_ch, 8, 3, padding=1)
        self.c2 = torch.nn.Conv2d(8, 1, 3, padding=1)

    def forward(self, xi, i):
        return self.c2(torch.tanh(self.c1(xi) + (i.float() / 24.0).view(-1, 1, 1, 1)))


@pytest.mark.parametrize("kind", ["prior", "replacement", "replacement_corrector", "amortized", "amortized_corrector"])
def test_arbitrary_python_eps_network_runs_stepwise(pkg, cuda, kind):
    """Any callable eps_model(xi, i) (sde_diffusion.py:11): a torch module the engine knows nothing about drives the chain
    through cfm_ddpm_step; same injected noise through the oracle's restatement of sampling.py -> fp32 agreement."""
    torch.manual_seed(3)
    amort = kind.startswith("amortized")
    net = TorchEps(2 if amort else 1).eval()
    n_corr = 2 if kind.endswith("corrector") else 0
    xT, cond = chain_inputs(200)
    Ns = CHAIN_NS
    slots = torch.randn(Ns, 2 + n_corr, *xT.shape, generator=torch.Generator().manual_seed(9))
    ddpm = pkg.DDPM(Ns)
    lik = pkg.InPainting(6, -2.0)
    net_gpu = TorchEps(2 if amort else 1).to(cuda).eval()
    net_gpu.load_state_dict(net.state_dict())
    eps_gpu = lambda xi, i: net_gpu(xi, i)
    feed = {"i": None, "k": 0}

    def oracle_noise(shape):     # sequential stream in the oracle's call order == slot order of one step
        z = slots[feed["i"], feed["k"]]
        feed["k"] += 1
        return z

    def eps_cpu(xi, t):          # the oracle passes t = i / Ns; the torch net takes i
        feed_i = torch.round(t * Ns).long()
        return net(xi, feed_i)

    tb = D.ddpm_tables(Ns)
    with torch.no_grad():
        if kind == "prior":
            got = pkg.get_prior_sample_fn(eps_gpu, ddpm, None, None, noise=slots.to(cuda))(xT.to(cuda))
            xi = xT
            for i in reversed(range(Ns)):
                eps = eps_cpu(xi, D.eps_time(i, Ns).repeat(2))
                xi, _ = D.posterior_step(tb, xi, eps, i, slots[i, 1] if i > 0 else 0.0)
            want = torch.clip(xi, -1, 1)
        elif not amort:
            cnd = pkg.Replacement(n_corrector=n_corr, delta=0.1)
            got = pkg.get_conditional_sample_fn(eps_gpu, ddpm, cnd, lik, noise=slots.to(cuda))(xT.to(cuda), cond.to(cuda))
            xi = xT
            for i in reversed(range(Ns)):
                nc = tb["sqrt_alphas_cumprod"][i] * cond + tb["sqrt_one_minus_alphas_cumprod"][i] * slots[i, 0]
                xi = torch.where(cond == -2.0, xi, nc)
                eps = eps_cpu(xi, D.eps_time(i, Ns).repeat(2))
                xi, _ = D.posterior_step(tb, xi, eps, i, slots[i, 1] if i > 0 else 0.0)
                for c in range(n_corr):
                    feed["i"], feed["k"] = i, 2 + c
                    xi = D.corrector_step(tb, eps_cpu, Ns, xi, i, 0.1, oracle_noise)
            want = torch.clip(xi, -1, 1)
        else:
            cnd = pkg.Amortized(n_corrector=n_corr, delta=0.2)
            got = pkg.get_conditional_sample_fn(eps_gpu, ddpm, cnd, lik, noise=slots.to(cuda))(xT.to(cuda), cond.to(cuda))
            xi = xT
            none_model = lambda x, t: eps_cpu(torch.cat([x, torch.full_like(x, -2.0)], dim=-3), t)
            for i in reversed(range(Ns)):
                eps = eps_cpu(torch.cat([xi, cond], dim=-3), D.eps_time(i, Ns).repeat(2))
                xi, _ = D.posterior_step(tb, xi, eps, i, slots[i, 1] if i > 0 else 0.0)
                for c in range(n_corr):
                    feed["i"], feed["k"] = i, 2 + c
                    xi = D.corrector_step(tb, none_model, Ns, xi, i, 0.2, oracle_noise)
            want = torch.clip(xi, -1, 1)
    r = rel_l2(got.cpu(), want)
    print(f"stepwise[{kind}] rel-L2 vs oracle = {r:.2e}")
    assert r < 1e-4


def test_stepwise_entry_point_equals_fused_chain(pkg, cuda):
    """cfm_ddpm_step driven with the engine's own forward as the network == cfm_sample_ddpm, bit for bit (same tables,
    noise layout, Philox streams): the fused chain is the stepwise chain without the host in the loop."""
    cfg, params, net = chain_net(pkg, cuda, 1)
    ddpm = pkg.DDPM(CHAIN_NS)
    xT, cond = chain_inputs(104)
    from_engine = lambda xi, i: net(xi, 1.0 * i / ddpm.Ns)
    cnd, lik = pkg.Replacement(n_corrector=1, delta=0.1), pkg.InPainting(6, -2.0)
    fused = pkg.get_conditional_sample_fn(pkg.EpsModel(net, ddpm), ddpm, cnd, lik, seed=77)(xT.to(cuda), cond.to(cuda))
    step = pkg.diffusion._stepwise_chain(from_engine, ddpm, xT.to(cuda), "replacement", condition=cond.to(cuda), pad_value=-2.0,
                                         replace_below_step=CHAIN_NS, noise_condition=True, seed=77, n_corrector=1,
                                         corrector_delta=0.1)
    assert torch.equal(fused, step)


def test_noise_advances_on_every_call_and_seed_pins_it(pkg, cuda):
    """the reference draws fresh randn_like noise per call.  Two successive calls of one sample() with
    the same xT must differ; an explicit seed= reproduces; the device normals have mean 0 / variance 1."""
    cfg, params, net = chain_net(pkg, cuda, 1)
    ddpm = pkg.DDPM(CHAIN_NS)
    xT, cond = chain_inputs(105)
    lik, cnd = pkg.InPainting(6, -2.0), pkg.Replacement()
    fn = pkg.get_conditional_sample_fn(pkg.EpsModel(net, ddpm), ddpm, cnd, lik)
    a, b = fn(xT.to(cuda), cond.to(cuda)), fn(xT.to(cuda), cond.to(cuda))
    assert not torch.equal(a, b)
    assert float((a - b).abs().mean()) > 1e-3                        # independent noise, not a rounding difference
    pinned = pkg.get_conditional_sample_fn(pkg.EpsModel(net, ddpm), ddpm, cnd, lik, seed=9)
    assert torch.equal(pinned(xT.to(cuda), cond.to(cuda)), pinned(xT.to(cuda), cond.to(cuda)))
    torch.manual_seed(1); c1 = fn(xT.to(cuda), cond.to(cuda))
    torch.manual_seed(1); c2 = fn(xT.to(cuda), cond.to(cuda))
    assert torch.equal(c1, c2)                                       # seeds come from torch's global generator
    # statistics of the device generator through the SDE step kernel: x = 0 + 0*dt + 1 * sqrt(1) * z
    lib = pkg._lib.load()
    z = torch.zeros(1 << 20, device=cuda)
    zero = torch.zeros_like(z)
    assert lib.cfm_sde_em_step(C.c_void_p(z.data_ptr()), C.c_void_p(zero.data_ptr()), None, 1.0, 1.0, None, C.c_uint64(1234), 7,
                               z.numel(), None) == 0
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.var()) - 1.0) < 1e-2
    z2 = torch.zeros(1 << 20, device=cuda)
    lib.cfm_sde_em_step(C.c_void_p(z2.data_ptr()), C.c_void_p(zero.data_ptr()), None, 1.0, 1.0, None, C.c_uint64(1234), 8, z2.numel(), None)
    assert abs(float((z * z2).mean())) < 5e-3                         # different streams are uncorrelated


def test_graph_cache_is_bounded_and_not_keyed_by_buffers(pkg, cuda):
    """a new noise tensor, seed or trajectory buffer must not instantiate a new graph."""
    cfg, params, net = chain_net(pkg, cuda, 1)
    ddpm = pkg.DDPM(8)
    e = net.engine()
    xT = torch.randn(2, 1, 16, 16, device=cuda)
    outs = []
    for seed in range(5):
        noise = torch.randn(8, 2, xT.numel(), device=cuda)          # a new tensor (new pointer) every call
        outs.append(e.sample_ddpm(xT, ddpm.tables(), mode="prior", noise=noise, use_graph=True))
        e.sample_ddpm(xT, ddpm.tables(), mode="prior", seed=seed, use_graph=True)
    assert e.cached_graphs == 1
    a = e.sample_ddpm(xT, ddpm.tables(), mode="prior", seed=3, use_graph=True)
    b = e.sample_ddpm(xT, ddpm.tables(), mode="prior", seed=3, use_graph=False)
    assert torch.equal(a, b)                                         # the replayed graph reads the CURRENT seed
    cfm = pkg.UNetModelWrapper(dim=(3, 16, 16), num_channels=32, num_res_blocks=1, channel_mult=(1, 2),
                               attention_resolutions="8", num_heads=2, precision="fp32").to(cuda).eval()
    x0 = torch.randn(2, 3, 16, 16, device=cuda)
    ts, dts = pkg.euler_time_grid(torch.linspace(0, 1, 4))
    trajs = [cfm.engine().sample_euler(x0, ts, dts, return_trajectory=True, use_graph=True)[1] for _ in range(3)]
    assert cfm.engine().cached_graphs == 1 and all(torch.equal(trajs[0], t) for t in trajs[1:])
    for B in range(1, 14):                                           # many shapes: the cache stays bounded (LRU)
        cfm.engine().sample_euler(torch.randn(B, 3, 16, 16, device=cuda), ts, dts, use_graph=True)
    assert cfm.engine().cached_graphs <= 8


def test_em_step_matches_reference_golden(pkg, cuda):
    g = np.load(os.path.join(GOLD, "sde_steps.npz"))
    x, eps, z = (torch.from_numpy(g[k]).to(cuda) for k in ("em.x", "em.eps", "em.z"))
    ddpm = pkg.DDPM(1000)
    for i in (0, 1, 500, 999):
        got = ddpm.em_step(x, eps, i, z=z).cpu()
        want = torch.from_numpy(g[f"em.out{i}"])
        err = float((got - want).abs().max())
        print(f"em_step[{i}] max abs diff vs reference = {err:.2e}")
        assert err <= 4e-6 * float(want.abs().max())                # fp32: the exp / division round differently at most


@pytest.mark.parametrize("with_score", [True, False])
def test_sde_sampler_matches_oracle(pkg, cuda, with_score):
    """SF2M sampling (conditional_mnist.ipynb cells 11-12): two class-conditional U-Nets, Euler-Maruyama, dt = 0.05."""
    kw = dict(class_cond=True, num_classes=10)
    cfg = O.config_from_wrapper((1, 16, 16), 32, 1, channel_mult=(1, 2), attention_resolutions="8", **kw)
    pd, ps = O.seeded_params(cfg, 61), O.seeded_params(cfg, 62)
    mk = lambda: pkg.UNetModelWrapper(dim=(1, 16, 16), num_channels=32, num_res_blocks=1, channel_mult=(1, 2),
                                      attention_resolutions="8", precision="fp32", **kw)
    md, ms = mk(), mk()
    md.load_state_dict(pd); ms.load_state_dict(ps)
    md, ms = md.to(cuda).eval(), ms.to(cuda).eval()
    x0 = torch.randn(4, 1, 16, 16, generator=torch.Generator().manual_seed(1))
    y = torch.tensor([0, 3, 7, 9])
    ts = torch.linspace(0, 1, 2)
    n_steps = 20
    noise = torch.randn(n_steps, *x0.shape, generator=torch.Generator().manual_seed(2))
    it = iter(noise)
    drift = lambda t, x: O.wrapper_forward(cfg, pd, t, x, y)
    score = (lambda t, x: O.wrapper_forward(cfg, ps, t, x, y)) if with_score else None
    want = I.sde_euler_maruyama(drift, score, x0, ts, 0.05, 0.1, lambda s: next(it))
    got = pkg.sample_sde(md, ms if with_score else None, x0.to(cuda), ts, 0.05, sigma=0.1, y=y.to(cuda), noise=noise.to(cuda))
    r = rel_l2(got.cpu(), want)
    print(f"sde[score={with_score}] rel-L2 vs oracle = {r:.2e}")
    assert r < 2e-4
    # generic callables take the stepwise kernel: same result as the fused native loop
    gen = pkg.sample_sde(lambda t, x, yy: md(t, x, yy), (lambda t, x, yy: ms(t, x, yy)) if with_score else None, x0.to(cuda), ts, 0.05,
                         sigma=0.1, y=y.to(cuda), noise=noise.to(cuda))
    assert torch.equal(gen, got)
    a = pkg.sample_sde(md, None, x0.to(cuda), ts, 0.05, sigma=0.1, y=y.to(cuda), seed=5)
    b = pkg.sample_sde(md, None, x0.to(cuda), ts, 0.05, sigma=0.1, y=y.to(cuda), seed=6)
    assert torch.isfinite(a).all() and not torch.equal(a, b)


def test_bilinear_resize_and_hyperresolution(pkg, cuda):
    """cfm_resize_bilinear against ATen's CPU bilinear (align_corners=False) on down-, up- and odd-ratio resizes, and
    pkg.HyperResolution.sample / downsample_images against the reference's own outputs (ddpm_chains.npz lik.hyperres)."""
    g = np.load(os.path.join(GOLD, "ddpm_chains.npz"))
    imgs = torch.from_numpy(g["lik.images"])
    got = pkg.HyperResolution(7, 7).sample(imgs.to(cuda)).cpu()
    want = torch.from_numpy(g["lik.hyperres"])
    err = float((got - want).abs().max())
    print(f"HyperResolution.sample max abs diff vs reference = {err:.2e}")
    assert err < 1e-6
    rs = torch.Generator().manual_seed(0)
    for (B, Cc, H, W, h, w) in [(3, 3, 64, 64, 16, 16), (2, 1, 28, 28, 7, 7), (2, 3, 16, 16, 64, 64), (1, 2, 17, 23, 40, 9),
                                (4, 3, 32, 32, 128, 128), (1, 1, 5, 5, 5, 5)]:
        x = torch.randn(B, Cc, H, W, generator=rs)
        want = F.interpolate(x, size=(h, w), mode="bilinear", align_corners=False)
        got = pkg.resize_bilinear(x.to(cuda), (h, w)).cpu()
        assert got.shape == want.shape
        # fp32 tolerance: ATen's vectorised CPU kernel contracts the weighted sums into FMAs, this kernel rounds every
        # product (a few ulp at |x| ~ 3); the reference's own HyperResolution output above agrees to 6e-8
        assert float((got - want).abs().max()) < 4e-6, (B, Cc, H, W, h, w)
    assert torch.equal(pkg.downsample_images(imgs.to(cuda), (7, 7)).cpu(),
                       pkg.resize_bilinear(imgs.to(cuda), (7, 7)).cpu())
    with pytest.raises(RuntimeError):
        pkg.resize_bilinear(imgs, (7, 7))                            # CPU tensor: no fallback


def test_hyperresolution_amortized_chain(pkg, cuda):
    """The super-resolution DDPM configuration (config.py: hyperresolution + amortized): condition = bilinear down/up of
    the image, concatenated on channels; the chain runs natively from pkg.HyperResolution.sample's output."""
    cfg, params, net = chain_net(pkg, cuda, 2)
    ddpm = pkg.DDPM(CHAIN_NS)
    xT, _ = chain_inputs(301)
    imgs = torch.rand(2, 1, 16, 16, generator=torch.Generator().manual_seed(4)) * 2 - 1
    lik = pkg.HyperResolution(4, 4)
    condition = lik.sample(imgs.to(cuda))
    want_cond = D.hyperresolution_condition(imgs, (4, 4))
    assert float((condition.cpu() - want_cond).abs().max()) < 1e-6
    slots = torch.randn(CHAIN_NS, 2, *xT.shape, generator=torch.Generator().manual_seed(8))
    fn = pkg.get_conditional_sample_fn(lambda xi, i: net(xi, 1.0 * i / ddpm.Ns), ddpm, pkg.Amortized(), lik, noise=slots.to(cuda))
    got = fn(xT.to(cuda), condition)
    tb = D.ddpm_tables(CHAIN_NS)
    xi = xT
    for i in reversed(range(CHAIN_NS)):
        eps = O.unet_forward(cfg, params, torch.cat([xi, want_cond], dim=-3), D.eps_time(i, CHAIN_NS).repeat(2))
        xi, _ = D.posterior_step(tb, xi, eps, i, slots[i, 1] if i > 0 else 0.0)
    assert rel_l2(got.cpu(), torch.clip(xi, -1, 1)) < 2e-4


def test_fid_statistics_and_frechet_distance(pkg, cuda):
    """cfm_fid_accumulate (fp64 sums, batch by batch) + frechet_distance against the numpy / scipy restatement."""
    rs = np.random.RandomState(0)
    D_ = 96
    a = (rs.standard_normal((700, D_)) @ rs.standard_normal((D_, D_)) * 0.3 + rs.standard_normal(D_)).astype(np.float32)
    b = (rs.standard_normal((650, D_)) @ rs.standard_normal((D_, D_)) * 0.3 + 0.2).astype(np.float32)
    sa, sb = pkg.FIDStatistics(D_, device=cuda), pkg.FIDStatistics(D_, device=cuda)
    for lo in range(0, 700, 256):
        sa.update(torch.from_numpy(a[lo:lo + 256]).to(cuda))
    sb.update(torch.from_numpy(b).to(cuda))
    assert sa.n == 700 and sb.n == 650
    mu_a, cov_a = sa.mean_cov(); mu_b, cov_b = sb.mean_cov()
    wa, wb = D.fid_statistics(a), D.fid_statistics(b)
    assert np.allclose(mu_a.cpu().numpy(), wa[0], rtol=0, atol=1e-12) and np.allclose(cov_a.cpu().numpy(), wa[1], rtol=1e-10, atol=1e-11)
    assert np.allclose(cov_b.cpu().numpy(), wb[1], rtol=1e-10, atol=1e-11)
    got = pkg.frechet_distance(mu_a, cov_a, mu_b, cov_b)
    want = D.frechet_distance(wa[0], wa[1], wb[0], wb[1])
    print(f"FID: engine {got:.9f}  oracle {want:.9f}")
    assert abs(got - want) < 1e-6 * max(1.0, abs(want))
    # deterministic accumulation: two runs, same bits; a dimension that is not a multiple of the tile
    s1, s2 = pkg.FIDStatistics(50, device=cuda), pkg.FIDStatistics(50, device=cuda)
    f = torch.from_numpy(a[:333, :50].copy()).to(cuda)
    s1.update(f); s2.update(f)
    assert torch.equal(s1.outer, s2.outer) and torch.equal(s1.sum, s2.sum)
    assert np.allclose(s1.mean_cov()[1].cpu().numpy(), np.cov(a[:333, :50].astype(np.float64), rowvar=False), rtol=1e-10, atol=1e-11)
    # the compute_fid loop with a stand-in extractor
    gen = lambda _: torch.from_numpy((rs.uniform(0, 255, size=(64, 3, 8, 8))).astype(np.uint8)).to(cuda)
    feat = lambda im: im.float().flatten(1)[:, :D_] / 255.0
    v = pkg.compute_fid(gen, feat, num_gen=200, batch_size=64, reference=(mu_b, cov_b))
    assert np.isfinite(v) and v > 0


def test_error_norm_is_deterministic(pkg, cuda):
    """The dopri5 error norm: same inputs -> same bits on every launch (fixed-order block partials), and the value is the
    fp64 sum of the fp32 ratios squared."""
    n = 1024 * 3 * 32 * 32 + 5
    g = torch.Generator(device=cuda).manual_seed(0)
    y0, y1 = torch.randn(n, device=cuda, generator=g), torch.randn(n, device=cuda, generator=g)
    ks = [torch.randn(n, device=cuda, generator=g) for _ in range(7)]
    scratch = torch.zeros(1, dtype=torch.float64, device=cuda)
    vals = [float(pkg.rk_error_sumsq(y0, y1, ks, I._C_ERR, 0.01, 1e-5, 1e-5, scratch).item()) for _ in range(5)]
    assert len(set(vals)) == 1
    err = torch.zeros(n, device=cuda)
    first = True
    for c, k in zip(I._C_ERR, ks):
        if c != 0:
            err = torch.tensor(c, dtype=torch.float32) * k if first else err + torch.tensor(c, dtype=torch.float32, device=cuda) * k
            first = False
    ratio = (torch.tensor(0.01, device=cuda) * err) / (1e-5 + 1e-5 * torch.maximum(y0.abs(), y1.abs()))
    want = float(ratio.double().pow(2).sum())
    assert abs(vals[0] - want) < 1e-9 * want
    with pytest.raises(ValueError):
        pkg.rk_combine(y0.cpu(), [ks[0].cpu()], [1.0], 0.1)          # raw-pointer kernels: CPU tensors are rejected


def test_dopri5_dense_output_and_initial_step_kernels(pkg, cuda):
    """The interpolant and the initial-step norms of dopri5 as native kernels: the dense output is bit-identical to the
    oracle's expression (oracle/integrators.py, torchdiffeq `_interp_fit` / `_interp_evaluate`) evaluated by torch on the
    CPU in fp32; the scaled norm equals the fp64 sum of the fp32 ratios squared and is run-to-run deterministic."""
    n = 4 * 3 * 16 * 16 + 3
    g = torch.Generator().manual_seed(3)
    y, y1, ym, fa, fb = (torch.randn(n, generator=g) for _ in range(5))
    dt, xx = 0.0375, 0.6180339887
    coeffs = (2 * dt * (fb - fa) - 8 * (y1 + y) + 16 * ym,
              dt * (5 * fa - 3 * fb) + 18 * y + 14 * y1 - 32 * ym,
              dt * (fb - 4 * fa) - 11 * y - 5 * y1 + 16 * ym,
              dt * fa, y)
    a_, b_, c_, d_, e_ = coeffs
    want = (((a_ * xx + b_) * xx + c_) * xx + d_) * xx + e_
    got = pkg.rk_dense_output(y.to(cuda), y1.to(cuda), ym.to(cuda), fa.to(cuda), fb.to(cuda), dt, xx).cpu()
    assert torch.equal(got, want), float((got - want).abs().max())
    scratch = torch.zeros(1, dtype=torch.float64, device=cuda)
    for b in (None, fb):
        vals = [float(pkg.rk_scaled_sumsq(fa.to(cuda), None if b is None else b.to(cuda), y.to(cuda), 1e-4, 1e-5, scratch).item())
                for _ in range(3)]
        assert len(set(vals)) == 1
        num = fa if b is None else fa - b
        ref = float((num / (1e-5 + y.abs() * 1e-4)).double().pow(2).sum())
        assert abs(vals[0] - ref) <= 1e-12 * ref
    with pytest.raises(ValueError):
        pkg.rk_dense_output(y, y1, ym, fa, fb, dt, xx)               # CPU tensors are rejected


def test_input_validation(pkg, cuda):
    """shapes, devices and labels are checked before raw pointers reach the kernels."""
    cfg, params, net = chain_net(pkg, cuda, 1)
    ddpm = pkg.DDPM(8)
    e = net.engine()
    xT = torch.randn(3, 1, 16, 16, device=cuda)
    with pytest.raises(ValueError):
        e.sample_ddpm(xT, ddpm.tables(), mode="replacement", condition=torch.zeros(3, 2, 16, 16, device=cuda))
    out = e.sample_ddpm(xT, ddpm.tables(), mode="replacement", condition=torch.full((1, 1, 16, 16), -2.0, device=cuda), seed=1)
    assert out.shape == xT.shape                                      # a [1, C, H, W] condition broadcasts like torch.where
    m = pkg.UNetModelWrapper(dim=(1, 16, 16), num_channels=32, num_res_blocks=1, channel_mult=(1, 2), attention_resolutions="8",
                             class_cond=True, num_classes=4, precision="fp32").to(cuda).eval()
    x = torch.randn(2, 1, 16, 16, device=cuda)
    with pytest.raises(IndexError):
        m(torch.tensor(0.5, device=cuda), x, torch.tensor([0, 4], device=cuda))
    dev_before = torch.cuda.current_device()
    pkg.UNetModel(image_size=16, in_channels=1, model_channels=32, out_channels=1, num_res_blocks=1, attention_resolutions=(2,),
                  channel_mult=(1, 2), precision="fp32").to(cuda).eval()(x, torch.rand(2, device=cuda))
    assert torch.cuda.current_device() == dev_before
