"""GPU parity of the sampling loops (Euler / dopri5 / DDPM reverse chains) against the CPU oracle,
with identical seeded weights, noise tensors and masks."""
import numpy as np
import pytest
import torch

from golden_configs import GOLDEN_CONFIGS
from oracle import ddpm as D
from oracle import integrators as I
from oracle import unet as O

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def small_cfm(pkg, cuda, precision, **kw):
    cfg = O.config_from_wrapper((3, 16, 16), 32, 1, channel_mult=(1, 2), attention_resolutions="8", num_heads=2, **kw)
    params = O.seeded_params(cfg, 31)
    m = pkg.UNetModelWrapper(dim=(3, 16, 16), num_channels=32, num_res_blocks=1, channel_mult=(1, 2),
                             attention_resolutions="8", num_heads=2, precision=precision, **kw)
    m.load_state_dict(params)
    return cfg, params, m.to(cuda).eval()


def test_step_kernels_bit_exact(pkg, cuda):
    """Euler update, uint8 conversion and RK stage combination are bit-identical to the torch CPU expressions."""
    x = torch.randn(5, 3, 16, 16); v = torch.randn(5, 3, 16, 16)
    got = pkg.rk_combine(x.to(cuda), [v.to(cuda)], [1.0], 0.01).cpu()
    assert torch.equal(got, x + torch.tensor(0.01) * v)
    ks = [torch.randn(1000) for _ in range(7)]
    coefs = I._C_MID
    want = torch.zeros(1000)
    first = True
    for c, k in zip(coefs, ks):
        if c != 0:
            want = torch.tensor(c, dtype=torch.float32) * k if first else want + torch.tensor(c, dtype=torch.float32) * k
            first = False
    y = torch.randn(1000)
    got = pkg.rk_combine(y.to(cuda), [k.to(cuda) for k in ks], coefs, 0.3).cpu()
    assert torch.equal(got, y + torch.tensor(0.3) * want)
    lib = pkg._lib.load()
    import ctypes as C
    z = (torch.randn(4097, device=cuda) * 2)
    out = torch.empty(4097, dtype=torch.uint8, device=cuda)
    assert lib.cfm_quantize_u8(C.c_void_p(out.data_ptr()), C.c_void_p(z.data_ptr()), 4097, None) == 0
    assert torch.equal(out.cpu(), D.to_uint8(z.cpu()))          # compute_fid.py:87


def test_box_condition_bit_exact(pkg, cuda):
    img = torch.rand(9, 3, 64, 64) * 2 - 1
    for cls, fn, patch in ((pkg.InPainting, D.inpainting_condition, 20), (pkg.OutPainting, D.outpainting_condition, 24)):
        lk = cls(patch_size=patch, pad_value=-2)
        torch.manual_seed(11)
        want = fn(img, patch, -2.0)
        torch.manual_seed(11)
        got = lk.sample(img.to(cuda)).cpu()
        assert torch.equal(got, want)
        assert torch.equal(got == -2.0, want == -2.0)
    with pytest.raises(RuntimeError):
        pkg.InPainting(patch_size=20, pad_value=-2).sample(torch.zeros(1, 1, 28, 28, device=cuda))


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 3e-2)])
def test_euler_trajectory_matches_oracle(pkg, cuda, precision, tol):
    cfg, params, m = small_cfm(pkg, cuda, precision)
    x0 = torch.randn(6, 3, 16, 16)
    t_span = torch.linspace(0, 1, 11)
    want = I.euler_trajectory(lambda t, x: O.wrapper_forward(cfg, params, t, x), x0, t_span)
    node = pkg.NeuralODE(m, solver="euler", sensitivity="adjoint")
    got = node.trajectory(x0.to(cuda), t_span.to(cuda)).cpu()
    assert got.shape == want.shape == (11, 6, 3, 16, 16)
    assert torch.equal(got[0], x0)
    drift = rel_l2(got[-1], want[-1])
    print(f"euler[{precision}] final-sample drift rel-L2 = {drift:.3e}")
    assert drift < tol
    # final-state-only entry point, with CUDA graph, agrees bit-for-bit with the trajectory's last state
    xf, img = pkg.sample_euler(m, x0.to(cuda), t_span, return_uint8=True, use_graph=True)
    assert torch.equal(xf.cpu(), got[-1])
    assert torch.equal(img.cpu(), D.to_uint8(got[-1]))
    xf2 = pkg.sample_euler(m, x0.to(cuda), t_span, use_graph=False)
    assert torch.equal(xf2.cpu(), got[-1])


def test_euler_conditional_state_drift(pkg, cuda):
    """utils_mnist2.py:118-134: state = cat(x, con), d(con)/dt = con  (SURVEY F8)."""
    cfg = O.config_from_wrapper((1, 28, 28), 32, 1, extra_in_channels=1)
    params = O.seeded_params(cfg, 32)
    m = pkg.InPaintModelWrapper(dim=(1, 28, 28), num_channels=32, num_res_blocks=1, precision="fp32")
    m.load_state_dict(params)
    m = m.to(cuda).eval()
    x0 = torch.randn(4, 1, 28, 28)
    torch.manual_seed(5)
    con = D.inpainting_condition(torch.rand(4, 1, 28, 28) * 2 - 1, 14)
    t_span = torch.linspace(0, 1, 6)

    def ode(t, s):
        v = O.inpaint_forward(cfg, params, s[:, :1], t, s[:, 1:])
        return torch.cat([v, s[:, 1:]], dim=1)
    want = I.euler_trajectory(ode, torch.cat([x0, con], 1), t_span)[-1]
    ts, dts = pkg.euler_time_grid(t_span)
    xf, _, _ = m.engine().sample_euler(x0.to(cuda), ts, dts, cond=con.to(cuda), cond_drift=True)
    assert rel_l2(xf.cpu(), want[:, :1]) < 2e-4
    # generic-callable path through NeuralODE (python vector field around the engine)
    node = pkg.NeuralODE(lambda t, s, args=None: torch.cat([m.forward(s[:, :1], t, con=s[:, 1:]), s[:, 1:]], 1), solver="euler")
    got = node.trajectory(torch.cat([x0, con], 1).to(cuda), t_span.to(cuda))[-1].cpu()
    assert rel_l2(got, want) < 2e-4


def test_dopri5_matches_oracle_controller(pkg, cuda):
    cfg, params, m = small_cfm(pkg, cuda, "fp32")
    x0 = torch.randn(4, 3, 16, 16)
    so, sg = {}, {}
    want = I.dopri5(lambda t, x: O.wrapper_forward(cfg, params, t, x), x0, [0.0, 1.0], 1e-4, 1e-4, stats=so)
    got = pkg.odeint(m, x0.to(cuda), torch.linspace(0, 1, 2), rtol=1e-4, atol=1e-4, method="dopri5", stats=sg)
    assert got.shape == (2, 4, 3, 16, 16)
    assert sg == so, (sg, so)                       # same accepted/rejected step sequence
    assert rel_l2(got[-1].cpu(), want[-1]) < 5e-4
    # tuple state (x, con) as in utils_mnist.py:96-108
    ca, cb = pkg.odeint(lambda t, s: (m(t, s[0]), s[1]), (x0.to(cuda), torch.ones(4, 2, device=cuda)),
                        torch.linspace(0, 1, 2), rtol=1e-4, atol=1e-4, method="dopri5")
    assert ca.shape == (2, 4, 3, 16, 16) and abs(float(cb[-1, 0, 0]) - np.e) < 1e-3


def test_dopri5_bf16_is_flagged_and_bounded(pkg, cuda):
    """the reference runs dopri5 at atol = rtol = 1e-4 on an fp32 net.  A bf16 model at that tolerance raises a
    RuntimeWarning (its rounding noise is above the tolerance); at a tolerance the precision supports it stays quiet, and
    its NFE count and result stay close to the fp32 run."""
    import warnings
    cfg, params, m32 = small_cfm(pkg, cuda, "fp32")
    _, _, m16 = small_cfm(pkg, cuda, "bf16")
    x0 = torch.randn(4, 3, 16, 16, device=cuda)
    t = torch.linspace(0, 1, 2)
    s32, s16 = {}, {}
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        ref = pkg.odeint(m32, x0, t, rtol=1e-4, atol=1e-4, method="dopri5", stats=s32)
        got = pkg.odeint(m16, x0, t, rtol=5e-3, atol=5e-3, method="dopri5", stats=s16)
    with pytest.warns(RuntimeWarning, match="bf16"):
        pkg.odeint(m16, x0, t, rtol=1e-4, atol=1e-4, method="dopri5")
    with pytest.warns(RuntimeWarning, match="bf16"):
        pkg.NeuralODE(m16, solver="dopri5", atol=1e-4, rtol=1e-4).trajectory(x0, t)
    assert s16["nfe"] <= s32["nfe"]
    assert rel_l2(got[-1].cpu(), ref[-1].cpu()) < 5e-2


def ddpm_setup(pkg, cuda, precision, in_ch, Ns):
    cfg = O.config_from_create_model(image_size=16, in_channels=in_ch, out_channels=1, num_channels=32, num_res_blocks=1,
                                     channel_mult="1,2", attention_resolutions="8", resblock_updown=True)
    params = O.seeded_params(cfg, 41)
    net = pkg.create_model(image_size=16, in_channels=in_ch, out_channels=1, num_channels=32, num_res_blocks=1,
                           channel_mult="1,2", attention_resolutions="8", resblock_updown=True, precision=precision)
    net.load_state_dict(params)
    net = net.to(cuda).eval()
    ddpm = pkg.DDPM(Ns)
    eps_oracle = lambda xi, t: O.unet_forward(cfg, params, xi, t)
    return net, ddpm, eps_oracle


class NoiseTape:
    """Replays one [Ns, 2, n] tensor in the oracle's call order (q_sample draw, then posterior draw)."""

    def __init__(self, Ns, shape, mode):
        self.t = torch.randn(Ns, 2, int(np.prod(shape)))
        self.i, self.slot, self.Ns, self.mode = Ns - 1, 0 if mode == "replacement" else 1, Ns, mode

    def __call__(self, shape):
        z = self.t[self.i, self.slot].reshape(shape)
        if self.mode == "replacement" and self.slot == 0:
            self.slot = 1
        else:
            self.i -= 1
            self.slot = 0 if self.mode == "replacement" else 1
        return z


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 6e-2)])
def test_ddpm_replacement_inpainting(pkg, cuda, precision, tol):
    Ns = 40
    net, ddpm, eps_oracle = ddpm_setup(pkg, cuda, precision, 1, Ns)
    torch.manual_seed(2)
    img = torch.rand(3, 1, 16, 16) * 2 - 1
    cond = img.clone(); cond[:, :, 5:11, 4:10] = -2.0
    xT = torch.randn(3, 1, 16, 16)
    tape = NoiseTape(Ns, xT.shape, "replacement")
    want = D.sample_replacement(eps_oracle, Ns, xT, cond, tape)
    fn = pkg.get_conditional_sample_fn(pkg.EpsModel(net, ddpm), ddpm, pkg.Replacement(start_fraction=1.0, noise=True),
                                       pkg.InPainting(6, -2.0), noise=tape.t)
    got = fn(xT.to(cuda), cond.to(cuda)).cpu()
    assert got.abs().max() <= 1.0
    d = rel_l2(got, want)
    print(f"ddpm replacement[{precision}] final-sample drift = {d:.3e}")
    assert d < tol
    # start_fraction < 1 and un-noised condition variants take the same code path as the oracle
    # (a noise-free chain through a random-weight net is chaotic: 1e-6 perturbations grow to 1e-3, so keep the noise)
    tape2 = NoiseTape(Ns, xT.shape, "amortized")     # noise=False: only the posterior draw is consumed
    want2 = D.sample_replacement(eps_oracle, Ns, xT, cond, tape2, start_fraction=0.5, noise_condition=False)
    fn2 = pkg.get_conditional_sample_fn(pkg.EpsModel(net, ddpm), ddpm, pkg.Replacement(start_fraction=0.5, noise=False),
                                        pkg.InPainting(6, -2.0), noise=tape2.t, use_graph=True)
    assert rel_l2(fn2(xT.to(cuda), cond.to(cuda)).cpu(), want2) < tol


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 6e-2)])
def test_ddpm_amortized_and_prior(pkg, cuda, precision, tol):
    Ns = 30
    net, ddpm, eps_oracle = ddpm_setup(pkg, cuda, precision, 2, Ns)
    cond = torch.rand(2, 1, 16, 16) * 2 - 1; cond[:, :, 3:9, 3:9] = -2.0
    xT = torch.randn(2, 1, 16, 16)
    tape = NoiseTape(Ns, xT.shape, "amortized")
    want = D.sample_amortized(eps_oracle, Ns, xT, cond, tape)
    fn = pkg.get_conditional_sample_fn(pkg.EpsModel(net, ddpm), ddpm, pkg.Amortized(0.9, 0, 0.1), pkg.InPainting(6, -2.0),
                                       noise=tape.t)
    assert rel_l2(fn(xT.to(cuda), cond.to(cuda)).cpu(), want) < tol
    # prior sampling with an amortised network substitutes none_like(x) = pad_value for the condition
    tape = NoiseTape(Ns, xT.shape, "amortized")
    want = D.sample_amortized(eps_oracle, Ns, xT, torch.full_like(cond, -2.0), tape)
    pf = pkg.get_prior_sample_fn(pkg.EpsModel(net, ddpm), ddpm, pkg.Amortized(0.9, 0, 0.1), pkg.InPainting(6, -2.0), noise=tape.t)
    assert rel_l2(pf(xT.to(cuda)).cpu(), want) < tol


def test_ddpm_device_rng_statistics(pkg, cuda):
    """Seeded Philox path: finite, clipped, reproducible, seed-sensitive."""
    net, ddpm, _ = ddpm_setup(pkg, cuda, "bf16", 1, 30)
    fn = lambda seed: pkg.get_prior_sample_fn(pkg.EpsModel(net, ddpm), ddpm, None, None, seed=seed)(torch.randn(4, 1, 16, 16, generator=torch.Generator().manual_seed(0)).to(cuda))
    a, b, c = fn(1), fn(1), fn(2)
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert torch.isfinite(a).all() and a.abs().max() <= 1.0


@pytest.mark.parametrize("precision,tol", [("fp32", 3e-4), ("bf16", 4e-2)])
def test_classifier_free_guidance_euler(pkg, cuda, precision, tol):
    """BASELINE config 3 (extension): two evaluations per step, v = v_c + w (v_c - v_u), the unconditional
    branch being the class-conditional model without its label embedding.  Checked against the oracle loop; w = 0
    must reproduce the plain conditional sampler bit for bit; use_graph replays the same captured step."""
    cfg, params, m = small_cfm(pkg, cuda, precision, class_cond=True, num_classes=10)
    x0 = torch.randn(6, 3, 16, 16)
    y = torch.tensor([0, 3, 9, 1, 1, 7])
    t_span = torch.linspace(0, 1, 5)
    w = 1.5
    want = I.euler_cfg_trajectory(lambda t, x: O.wrapper_forward(cfg, params, t, x, y),
                                  lambda t, x: O.wrapper_forward(cfg, params, t, x, None, drop_labels=True),
                                  w, x0, t_span)[-1]
    got = pkg.sample_euler(m, x0.to(cuda), t_span, y=y.to(cuda), guidance_weight=w, use_graph=False)
    r = rel_l2(got.cpu(), want)
    print(f"cfg[{precision}] rel-L2 = {r:.3e}")
    assert r < tol, r
    got_graph = pkg.sample_euler(m, x0.to(cuda), t_span, y=y.to(cuda), guidance_weight=w, use_graph=True)
    assert torch.equal(got, got_graph)
    plain = pkg.sample_euler(m, x0.to(cuda), t_span, y=y.to(cuda), use_graph=False)
    zero_w = pkg.sample_euler(m, x0.to(cuda), t_span, y=y.to(cuda), guidance_weight=0.0, use_graph=False)
    assert torch.equal(plain, zero_w)
    assert rel_l2(got.cpu(), plain.cpu()) > 1e-3           # guidance does change the trajectory
    _, _, unc = small_cfm(pkg, cuda, precision)
    with pytest.raises(Exception):
        pkg.sample_euler(unc, x0.to(cuda), t_span, guidance_weight=w)   # needs a class-conditional model


class SlotTape:
    """Replays a [Ns, 2 + n_corrector, n] noise tensor in the reference's call order for the Replacement sampler:
    per chain step i (descending): q_sample draw (slot 0), posterior draw (slot 1, skipped at i = 0), corrector draws."""

    def __init__(self, Ns, shape, n_corrector):
        self.t = torch.randn(Ns, 2 + n_corrector, int(np.prod(shape)))
        self.order = []
        for i in reversed(range(Ns)):
            self.order.append((i, 0))
            if i > 0:
                self.order.append((i, 1))
            self.order += [(i, 2 + c) for c in range(n_corrector)]
        self.pos = 0

    def __call__(self, shape):
        i, s = self.order[self.pos]
        self.pos += 1
        return self.t[i, s].reshape(shape)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 6e-2)])
def test_ddpm_replacement_with_langevin_corrector(pkg, cuda, precision, tol):
    """Predictor-corrector Replacement chain (sampling.py:241-256): one extra U-Net evaluation and one Langevin step
    per corrector, the mask blend of step i running before that step's U-Net call."""
    Ns, n_corr, delta = 24, 2, 0.1
    net, ddpm, eps_oracle = ddpm_setup(pkg, cuda, precision, 1, Ns)
    torch.manual_seed(4)
    img = torch.rand(3, 1, 16, 16) * 2 - 1
    cond = img.clone(); cond[:, :, 4:10, 6:12] = -2.0
    xT = torch.randn(3, 1, 16, 16)
    tape = SlotTape(Ns, xT.shape, n_corr)
    want = D.sample_replacement(eps_oracle, Ns, xT, cond, tape, n_corrector=n_corr, delta=delta)
    assert tape.pos == len(tape.order)
    for use_graph in (False, True):
        fn = pkg.get_conditional_sample_fn(pkg.EpsModel(net, ddpm), ddpm,
                                           pkg.Replacement(delta=delta, start_fraction=1.0, noise=True, n_corrector=n_corr),
                                           pkg.InPainting(6, -2.0), noise=tape.t, use_graph=use_graph)
        got = fn(xT.to(cuda), cond.to(cuda)).cpu()
        assert got.abs().max() <= 1.0
        d = rel_l2(got, want)
        print(f"ddpm replacement + {n_corr} correctors [{precision}, graph={use_graph}] drift = {d:.3e}")
        assert d < tol


class AmortTape(SlotTape):
    """Amortized chain: no q_sample draw; posterior draw (slot 1, skipped at i = 0), then the corrector draws."""

    def __init__(self, Ns, shape, n_corrector):
        super().__init__(Ns, shape, n_corrector)
        self.order = [(i, s) for (i, s) in self.order if s != 0]


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 6e-2)])
def test_ddpm_amortized_with_langevin_corrector(pkg, cuda, precision, tol):
    """Amortized predictor-corrector chain (sampling.py:113-127): the predictor's U-Net call sees the condition, the
    corrector's sees likelihood.none_like(xi) = pad_value everywhere (sampling.py:36-37)."""
    Ns, n_corr, delta = 24, 1, 0.2
    net, ddpm, eps_oracle = ddpm_setup(pkg, cuda, precision, 2, Ns)
    torch.manual_seed(9)
    img = torch.rand(3, 1, 16, 16) * 2 - 1
    cond = img.clone(); cond[:, :, 3:9, 5:11] = -2.0
    xT = torch.randn(3, 1, 16, 16)
    tape = AmortTape(Ns, xT.shape, n_corr)
    want = D.sample_amortized(eps_oracle, Ns, xT, cond, tape, n_corrector=n_corr, delta=delta, none_value=-2.0)
    assert tape.pos == len(tape.order)
    plain = D.sample_amortized(eps_oracle, Ns, xT, cond, AmortTape(Ns, xT.shape, 0))
    assert rel_l2(plain, want) > 10 * tol          # the corrector (and its none-condition) is visible in the result
    for use_graph in (False, True):
        fn = pkg.get_conditional_sample_fn(pkg.EpsModel(net, ddpm), ddpm, pkg.Amortized(0.9, n_corr, delta),
                                           pkg.InPainting(6, -2.0), noise=tape.t, use_graph=use_graph)
        got = fn(xT.to(cuda), cond.to(cuda)).cpu()
        assert got.abs().max() <= 1.0
        d = rel_l2(got, want)
        print(f"ddpm amortized + {n_corr} corrector [{precision}, graph={use_graph}] drift = {d:.3e}")
        assert d < tol


def _dopri5_shard_worker(rank, world, port, q):
    import os
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    pkg = ge.load_package()
    from oracle import unet as O2
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)     # both ranks share cuda:0; scalars travel over gloo
    cfg = O2.config_from_wrapper((3, 16, 16), 32, 1, channel_mult=(1, 2), attention_resolutions="8", num_heads=2)
    m = pkg.UNetModelWrapper(dim=(3, 16, 16), num_channels=32, num_res_blocks=1, channel_mult=(1, 2),
                             attention_resolutions="8", num_heads=2, precision="fp32")
    m.load_state_dict(O2.seeded_params(cfg, 31))
    m = m.to("cuda:0").eval()
    torch.manual_seed(11)
    x0 = torch.randn(6, 3, 16, 16)
    t = torch.linspace(0, 1, 2)
    stats = {}
    traj = pkg.odeint_sharded(m, x0, t, rtol=1e-4, atol=1e-4, stats=stats)
    lo, hi = pkg.shard_range(6, rank, world)
    q.put((rank, lo, hi, traj[-1].cpu(), stats))
    dist.destroy_process_group()


def test_dopri5_sharded_shares_one_step_controller(pkg, cuda):
    """Two ranks, each with half of the batch, all-reduce the error norms: both take exactly the accepted / rejected
    step sequence of a single process integrating the whole batch, and the final states agree with it."""
    import os
    import torch.multiprocessing as mp
    cfg, params, m = small_cfm(pkg, cuda, "fp32")
    torch.manual_seed(11)
    x0 = torch.randn(6, 3, 16, 16)
    t = torch.linspace(0, 1, 2)
    ref_stats = {}
    ref = pkg.odeint(m, x0.to(cuda), t.to(cuda), rtol=1e-4, atol=1e-4, method="dopri5", stats=ref_stats)[-1].cpu()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dopri5_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
    for rank, lo, hi, xf, stats in res:
        assert stats["steps"] == ref_stats["steps"] and stats["accepted"] == ref_stats["accepted"], (stats, ref_stats)
        assert rel_l2(xf, ref[lo:hi]) < 1e-5


def _reference_noise_tensor(case, Ns, shape):
    """The draws the reference's sampler makes after ``torch.manual_seed(case.seed)`` (q_sample draw while blending,
    posterior draw for i > 0, corrector draws), laid out as the engine's [Ns, 2 + n_corrector, n] tensor."""
    n_corr = case["n_corrector"]
    t = torch.zeros(Ns, 2 + n_corr, int(np.prod(shape)))
    repl = case["kind"] == "replacement" and not case["prior"]
    torch.manual_seed(case["seed"])
    for i in reversed(range(Ns)):
        if repl and case["noise"] and i < int(Ns * case["start_fraction"]):
            t[i, 0] = torch.randn(shape).flatten()
        if i > 0:
            t[i, 1] = torch.randn(shape).flatten()
        for c in range(n_corr):
            t[i, 2 + c] = torch.randn(shape).flatten()
    return t


@pytest.mark.parametrize("name", ["prior", "prior_amortized", "replacement", "replacement_raw_condition",
                                  "replacement_corrector2", "amortized", "amortized_corrector1"])
def test_ddpm_chains_match_reference_golden(pkg, cuda, name):
    """Engine chains against outputs of the reference's own sampling.py (tests/golden/ddpm_chains.npz)."""
    import os
    from golden_configs import CHAIN_CASES, CHAIN_NS, chain_inputs
    case = CHAIN_CASES[name]
    want = torch.from_numpy(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ddpm_chains.npz"))[name])
    net, ddpm, _ = ddpm_setup(pkg, cuda, "fp32", case["in_ch"], CHAIN_NS)
    xT, cond = chain_inputs(case["seed"])
    noise = _reference_noise_tensor(case, CHAIN_NS, xT.shape)
    lik = pkg.InPainting(6, -2.0)
    if case["kind"] == "amortized":
        cnd = pkg.Amortized(0.9, case["n_corrector"], case["delta"])
    else:
        cnd = pkg.Replacement(delta=case["delta"], start_fraction=case["start_fraction"], noise=case["noise"],
                              n_corrector=case["n_corrector"])
    for use_graph in (False, True):
        if case["prior"]:
            got = pkg.get_prior_sample_fn(pkg.EpsModel(net, ddpm), ddpm, cnd, lik, noise=noise, use_graph=use_graph)(xT.to(cuda))
        else:
            got = pkg.get_conditional_sample_fn(pkg.EpsModel(net, ddpm), ddpm, cnd, lik, noise=noise,
                                                use_graph=use_graph)(xT.to(cuda), cond.to(cuda))
        d = rel_l2(got.cpu(), want)
        print(f"{name} [graph={use_graph}] vs reference chain: rel-L2 = {d:.3e}")
        assert d < 1e-3
