"""Oracle pinning (CPU): against the committed golden vectors (outputs of the reference's own
code) and, when /root/reference is present, against the live vendored modules."""
import os

import numpy as np
import pytest
import torch

from _refload import load_reference, reference_available
from golden_configs import CHAIN_CASES, CHAIN_NS, GOLDEN_CONFIGS, N_RANDOM_CONFIGS, chain_inputs, chain_net_cfg, random_config
from oracle import ddpm as D
from oracle import unet as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", list(GOLDEN_CONFIGS))
def test_unet_oracle_matches_reference_golden(name):
    cfg, batch, seed = GOLDEN_CONFIGS[name]
    g = np.load(os.path.join(GOLD, f"unet_{name}.npz"))
    assert int(g["n_params"]) == O.count_params(cfg)
    p = O.seeded_params(cfg, int(g["seed"]))
    out = O.unet_forward(cfg, p, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]))
    want = torch.from_numpy(g["out"])
    rel = float((out - want).norm() / want.norm())
    assert rel < 1e-5, rel      # same ops, same machine class: essentially bit-identical


def test_known_parameter_counts():
    # SURVEY F4 / BASELINE.md: notebook prints 1.23M; CIFAR 35,746,307; flowers 68,156,163
    assert O.count_params(GOLDEN_CONFIGS["cifar"][0]) == 35_746_307
    assert O.count_params(GOLDEN_CONFIGS["flowers_ddpm"][0]) == 68_156_163
    c = O.config_from_create_model(image_size=28, in_channels=1, out_channels=1, num_channels=32, num_res_blocks=1,
                                   channel_mult="1, 2, 2", resblock_updown=True)
    assert O.count_params(c) == 1_225_185
    fl = O.flops_per_sample(GOLDEN_CONFIGS["cifar"][0])
    assert abs(fl["total"] / 1e9 - 12.444) < 0.01


def test_ddpm_tables_match_reference_golden():
    g = np.load(os.path.join(GOLD, "ddpm_reference.npz"))
    for Ns in (1000, 20):
        tb = D.ddpm_tables(Ns)
        for k in tb:
            if k == "ts":
                continue
            assert np.array_equal(tb[k].numpy(), g[f"Ns{Ns}.{k}"]), (Ns, k)
    tb = D.ddpm_tables(1000)
    assert abs(float(tb["alphas_cumprod"][0]) - 0.9998998) < 1e-7
    assert abs(float(tb["posterior_log_variance_clipped"][0]) + 46.0517) < 1e-3
    x, e, z = (torch.from_numpy(g[f"step.{k}"]) for k in ("x", "eps", "z"))
    for i in (0, 1, 500, 999):
        got, _ = D.posterior_step(tb, x, e, i, z)
        assert np.array_equal(got.numpy(), g[f"step.out{i}"]), i
        assert float(D.eps_time(i, 1000)) == float(g[f"step.t{i}"][0])


@pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted")
def test_oracle_matches_live_reference_unet():
    ref = load_reference()
    cfg, _, _ = GOLDEN_CONFIGS["tiny_neworder"]
    m = ref.unet.UNetModel(image_size=cfg.image_size, in_channels=cfg.in_channels, model_channels=cfg.model_channels,
                           out_channels=cfg.out_channels, num_res_blocks=cfg.num_res_blocks,
                           attention_resolutions=cfg.attention_ds, channel_mult=cfg.channel_mult,
                           num_heads=cfg.num_heads, use_new_attention_order=True).eval()
    assert list(m.state_dict().keys()) == list(O.param_shapes(cfg).keys())
    p = O.seeded_params(cfg, 11)
    m.load_state_dict(p)
    x, t = torch.randn(2, 3, 16, 16), torch.rand(2)
    with torch.no_grad():
        want = m(x, t)
    assert torch.equal(want, O.unet_forward(cfg, p, x, t))


@pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted")
@pytest.mark.parametrize("seed", range(N_RANDOM_CONFIGS))
def test_oracle_matches_live_reference_on_random_configs(seed):
    """The oracle pinned across the constructor's keyword space: for every config of the seeded sweep the GPU tests use
    (golden_configs.random_config), the vendored UNetModel and the oracle agree bit for bit - parameter names included."""
    ref = load_reference()
    kw, B = random_config(seed)
    kw.pop("num_classes", None)     # the vendored forward(x, timesteps) takes no labels (unet.py:708); the label path is torchcfm's
    cfg = O.config_from_create_model(**kw)
    m = ref.unet.UNetModel(image_size=cfg.image_size, in_channels=cfg.in_channels, model_channels=cfg.model_channels,
                           out_channels=cfg.out_channels, num_res_blocks=cfg.num_res_blocks,
                           attention_resolutions=cfg.attention_ds, channel_mult=cfg.channel_mult,
                           num_heads=cfg.num_heads, num_head_channels=cfg.num_head_channels,
                           num_heads_upsample=cfg.num_heads_upsample, use_scale_shift_norm=cfg.use_scale_shift_norm,
                           resblock_updown=cfg.resblock_updown, use_new_attention_order=cfg.use_new_attention_order).eval()
    assert list(m.state_dict().keys()) == list(O.param_shapes(cfg).keys())
    p = O.seeded_params(cfg, 900 + seed)
    m.load_state_dict(p)
    rs = np.random.RandomState(seed)
    x = torch.from_numpy(rs.standard_normal((B, kw["in_channels"], kw["image_size"], kw["image_size"])).astype(np.float32))
    t = torch.from_numpy(rs.uniform(0, 1, size=(B,)).astype(np.float32))
    with torch.no_grad():
        want = m(x, t)
    assert torch.equal(want, O.unet_forward(cfg, p, x, t)), kw


def test_mask_sampler_rng_order_and_bounds():
    # h is drawn before w, both randint(5, S-p-5) from the global CPU generator (likelihoods.py:49-53)
    img = torch.rand(5, 1, 28, 28) * 2 - 1
    boxes = []
    torch.manual_seed(3)
    cond = D.inpainting_condition(img, 14, -2.0, boxes=boxes)
    torch.manual_seed(3)
    for (h, w) in boxes:
        assert h == int(torch.randint(5, 28 - 14 - 5, size=())) and w == int(torch.randint(5, 28 - 14 - 5, size=()))
        assert 5 <= h < 9 and 5 <= w < 9
    for k, (h, w) in enumerate(boxes):
        m = cond[k] == -2.0
        assert int(m.sum()) == 14 * 14 and bool(m[:, h:h + 14, w:w + 14].all())
        assert torch.equal(cond[k][~m], img[k][~m])
    with pytest.raises(RuntimeError):   # MNIST with the default patch 20 cannot draw a box
        D.inpainting_condition(img, 20)
    out = D.outpainting_condition(img, 14, -2.0)
    assert int((out != -2.0).sum()) == 5 * 14 * 14


def test_samplers_consume_noise_in_reference_order():
    calls = []

    def noise(shape):
        calls.append(tuple(shape))
        return torch.zeros(shape)

    eps = lambda x, t: torch.zeros(x.shape[0], 1, 8, 8)
    xT = torch.randn(2, 1, 8, 8)
    cond = torch.full((2, 1, 8, 8), -2.0); cond[:, :, :4] = 0.3
    D.sample_replacement(eps, 30, xT, cond, noise)   # Ns > 20 keeps beta < 1
    assert len(calls) == 30 + 29          # q_sample every step, posterior noise for i > 0
    calls.clear()
    out = D.sample_amortized(lambda x, t: torch.zeros(x.shape[0], 1, 8, 8), 30, xT, cond, noise)
    assert len(calls) == 29 and out.abs().max() <= 1


@pytest.mark.parametrize("name", list(CHAIN_CASES))
def test_ddpm_chains_match_reference_sampling(name):
    """Whole reverse chains of the oracle against outputs of the reference's own ``sampling.py`` (ddpm_chains.npz,
    tests/golden/make_golden.py): prior / Replacement / Amortized, with and without Langevin correctors.  The reference
    draws with ``torch.randn_like`` from the global generator; the oracle's ``noise`` callable draws ``torch.randn`` of
    the same shapes in the same order after the same seed, so the two chains see identical noise."""
    case = CHAIN_CASES[name]
    want = torch.from_numpy(np.load(os.path.join(GOLD, "ddpm_chains.npz"))[name])
    cfg = chain_net_cfg(case["in_ch"])
    params = O.seeded_params(cfg, 41)
    eps = lambda xi, t: O.unet_forward(cfg, params, xi, t)
    xT, cond = chain_inputs(case["seed"])
    noise = lambda shape: torch.randn(tuple(shape))
    torch.manual_seed(case["seed"])
    if case["prior"] and case["kind"] == "amortized":
        got = D.sample_amortized(eps, CHAIN_NS, xT, torch.full_like(xT, -2.0), noise)     # none_like as the condition
    elif case["prior"]:
        got = D.sample_prior(eps, CHAIN_NS, xT, noise)
    elif case["kind"] == "amortized":
        got = D.sample_amortized(eps, CHAIN_NS, xT, cond, noise, n_corrector=case["n_corrector"], delta=case["delta"],
                                 none_value=-2.0)
    else:
        got = D.sample_replacement(eps, CHAIN_NS, xT, cond, noise, pad_value=-2.0, start_fraction=case["start_fraction"],
                                   noise_condition=case["noise"], n_corrector=case["n_corrector"], delta=case["delta"])
    err = float((got - want).abs().max())
    print(f"{name}: max abs diff vs reference chain = {err:.2e}")
    assert err < 2e-4


def test_condition_builders_match_reference_likelihoods():
    """likelihoods.py In/OutPainting/HyperResolution ``sample`` outputs (ddpm_chains.npz, ``lik.*``) against the oracle's
    builders drawing from the same seeded global generator."""
    g = np.load(os.path.join(GOLD, "ddpm_chains.npz"))
    imgs = torch.from_numpy(g["lik.images"])
    torch.manual_seed(7)
    assert np.array_equal(D.inpainting_condition(imgs, 14, -2.0).numpy(), g["lik.inpaint"])
    torch.manual_seed(8)
    assert np.array_equal(D.outpainting_condition(imgs, 10, -2.0).numpy(), g["lik.outpaint"])
    assert np.array_equal(D.hyperresolution_condition(imgs, (7, 7)).numpy(), g["lik.hyperres"])


def test_em_step_matches_reference_methods():
    """oracle.ddpm.em_step against the body of sampling.py:100-111 evaluated with the reference's own DDPM methods
    (tests/golden/sde_steps.npz, generated by make_golden.py sde): bit for bit."""
    g = np.load(os.path.join(GOLD, "sde_steps.npz"))
    x, eps, z = (torch.from_numpy(g[k]) for k in ("em.x", "em.eps", "em.z"))
    for i in (0, 1, 500, 999):
        got = D.em_step(1000, x, eps, i, z)
        assert np.array_equal(got.numpy(), g[f"em.out{i}"]), i


def test_fid_oracle_known_answers():
    """Frechet distance restatement: identical statistics -> 0; a pure mean shift -> |d|^2; commuting diagonal
    covariances -> closed form sum (sqrt(a) - sqrt(b))^2."""
    rs = np.random.RandomState(0)
    f = rs.standard_normal((500, 6))
    mu, cov = D.fid_statistics(f)
    assert abs(D.frechet_distance(mu, cov, mu, cov)) < 1e-8
    d = rs.standard_normal(6)
    assert abs(D.frechet_distance(mu + d, cov, mu, cov) - d.dot(d)) < 1e-8
    a, b = rs.uniform(0.5, 2, 6), rs.uniform(0.5, 2, 6)
    want = float(((np.sqrt(a) - np.sqrt(b)) ** 2).sum())
    assert abs(D.frechet_distance(mu, np.diag(a), mu, np.diag(b)) - want) < 1e-8


def test_sde_euler_maruyama_closed_form():
    """Oracle Euler-Maruyama on dx = -x dt + sigma dW with the noise switched off reduces to Euler on dx = -x dt:
    x_n = x_0 (1 - dt)^n; with noise on, the injected normals enter as sigma * sqrt(dt) * z."""
    from oracle import integrators as I
    x0 = torch.ones(3, 2)
    out = I.sde_euler_maruyama(lambda t, x: -x, None, x0, torch.tensor([0.0, 1.0]), 0.01, 0.0, lambda s: torch.zeros(s))
    assert torch.allclose(out, x0 * (1 - 0.01) ** 100, atol=1e-5)
    out = I.sde_euler_maruyama(lambda t, x: 0 * x, lambda t, x: 0 * x, x0, torch.tensor([0.0, 0.02]), 0.01, 0.5, lambda s: torch.ones(s))
    assert torch.allclose(out, x0 + 2 * 0.5 * 0.1, atol=1e-6)
