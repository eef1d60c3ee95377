"""Generate tests/golden/*.npz from the REFERENCE's own code (build container only).

Run:  python tests/golden/make_golden.py
Needs /root/reference.  The vendored ``UNetModel`` (AD/image_diffusion/unet.py) and ``DDPM``
(AD/image_diffusion/sde_diffusion.py) are imported by path, loaded with the seeded weights of
``oracle.unet.seeded_params`` (numpy legacy RandomState: frozen stream) and evaluated on seeded
inputs.  Only inputs/outputs are stored; the weights are re-generated from the seed at test time.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from _refload import load_reference, load_reference_samplers  # noqa: E402
from golden_configs import CHAIN_CASES, CHAIN_NS, chain_inputs, chain_net_cfg  # noqa: E402
from oracle import unet as O  # noqa: E402
from golden_configs import GOLDEN_CONFIGS  # noqa: E402


def _ref_unet(ref, cfg, params):
    model = ref.unet.UNetModel(
        image_size=cfg.image_size, in_channels=cfg.in_channels, model_channels=cfg.model_channels,
        out_channels=cfg.out_channels, num_res_blocks=cfg.num_res_blocks, attention_resolutions=cfg.attention_ds,
        channel_mult=cfg.channel_mult, num_classes=None, num_heads=cfg.num_heads,
        num_head_channels=cfg.num_head_channels, num_heads_upsample=cfg.num_heads_upsample,
        use_scale_shift_norm=cfg.use_scale_shift_norm, resblock_updown=cfg.resblock_updown,
        use_new_attention_order=cfg.use_new_attention_order).eval()
    model.load_state_dict(params)
    return model


def chains():
    """Whole reverse chains from the reference's own ``sampling.py`` (prior / Replacement / Amortized, with and without
    Langevin correctors) on a small seeded network; the global torch generator is seeded right before each call."""
    ref = load_reference_samplers()
    d = {}
    ddpm = ref.sde_diffusion.DDPM(CHAIN_NS)
    lik = ref.likelihoods.InPainting(patch_size=6, pad_value=-2.0)
    for name, case in CHAIN_CASES.items():
        cfg = chain_net_cfg(case["in_ch"])
        net = _ref_unet(ref, cfg, O.seeded_params(cfg, 41))
        eps_model = lambda xi, i, net=net: net(xi, 1.0 * i / ddpm.Ns)      # experiments/main.py:140
        if case["kind"] == "amortized":
            cnd = ref.conditioning.Amortized(p_cond=0.9, n_corrector=case["n_corrector"], delta=case["delta"])
        else:
            cnd = ref.conditioning.Replacement(delta=case["delta"], start_fraction=case["start_fraction"],
                                               noise=case["noise"], n_corrector=case["n_corrector"])
        xT, cond = chain_inputs(case["seed"])
        torch.manual_seed(case["seed"])
        if case["prior"]:
            out = ref.sampling.get_prior_sample_fn(eps_model, ddpm, cnd, lik)(xT.clone())
        else:
            out = ref.sampling.get_conditional_sample_fn(eps_model, ddpm, cnd, lik)(xT.clone(), cond.clone())
        d[name] = out.numpy()
        print(f"chain {name}: rms={float(out.pow(2).mean().sqrt()):.4f}")
    # likelihoods.py condition builders under a seeded global generator (box draws: h first, then w, per sample)
    rs = np.random.RandomState(77)
    imgs = torch.from_numpy(rs.uniform(-1, 1, size=(3, 1, 28, 28)).astype(np.float32))
    d["lik.images"] = imgs.numpy()
    torch.manual_seed(7)
    d["lik.inpaint"] = ref.likelihoods.InPainting(14, -2.0).sample(imgs).numpy()
    torch.manual_seed(8)
    d["lik.outpaint"] = ref.likelihoods.OutPainting(10, -2.0).sample(imgs).numpy()
    d["lik.hyperres"] = ref.likelihoods.HyperResolution(7, 7).sample(imgs).numpy()
    np.savez_compressed(os.path.join(HERE, "ddpm_chains.npz"), **d)


def sde_steps():
    """``em_step`` of the Amortized sampler (sampling.py:100-111) is a closure that the reference never calls; its body is
    evaluated here with the reference's own DDPM methods (backward_drift / backward_diffusion / score_from_noise)."""
    ref = load_reference()
    m = ref.sde_diffusion.DDPM(1000)
    rs = np.random.RandomState(7)
    xi = torch.from_numpy(rs.standard_normal((4, 3, 8, 8)).astype(np.float32))
    noise_hat = torch.from_numpy(rs.standard_normal((4, 3, 8, 8)).astype(np.float32))
    z = torch.from_numpy(rs.standard_normal((4, 3, 8, 8)).astype(np.float32))
    d = {"em.x": xi.numpy(), "em.eps": noise_hat.numpy(), "em.z": z.numpy()}
    for i in (0, 1, 500, 999):
        batched_times = torch.full((4,), i, dtype=torch.long)
        score_fn = m.score_from_noise
        drift = m.backward_drift(score_fn, xi, noise_hat, batched_times)
        diffusion = m.backward_diffusion(batched_times)
        dt = 1 / m.Ns
        x = xi - dt * drift + diffusion.unsqueeze(1).unsqueeze(2).unsqueeze(3) * z * np.sqrt(dt)
        d[f"em.out{i}"] = x.numpy()
    np.savez_compressed(os.path.join(HERE, "sde_steps.npz"), **d)
    print("em_step golden written")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "sde":
        return sde_steps()
    chains()
    sde_steps()
    ref = load_reference()
    assert ref is not None, "/root/reference is required to generate golden vectors"
    torch.set_num_threads(os.cpu_count())
    for name, (cfg, batch, seed) in GOLDEN_CONFIGS.items():
        model = ref.unet.UNetModel(
            image_size=cfg.image_size, in_channels=cfg.in_channels, model_channels=cfg.model_channels,
            out_channels=cfg.out_channels, num_res_blocks=cfg.num_res_blocks, attention_resolutions=cfg.attention_ds,
            channel_mult=cfg.channel_mult, num_classes=None, num_heads=cfg.num_heads,
            num_head_channels=cfg.num_head_channels, num_heads_upsample=cfg.num_heads_upsample,
            use_scale_shift_norm=cfg.use_scale_shift_norm, resblock_updown=cfg.resblock_updown,
            use_new_attention_order=cfg.use_new_attention_order).eval()
        params = O.seeded_params(cfg, seed)
        model.load_state_dict(params)
        rs = np.random.RandomState(1000 + seed)
        x = torch.from_numpy(rs.standard_normal((batch, cfg.in_channels, cfg.image_size, cfg.image_size)).astype(np.float32))
        t = torch.from_numpy(rs.uniform(0, 1, size=(batch,)).astype(np.float32))
        with torch.no_grad():
            out = model(x, t)
        n_params = sum(p.numel() for p in model.parameters())
        np.savez_compressed(os.path.join(HERE, f"unet_{name}.npz"), x=x.numpy(), t=t.numpy(), out=out.numpy(),
                            n_params=np.int64(n_params), seed=np.int64(seed))
        print(f"{name}: params={n_params} out_rms={float(out.pow(2).mean().sqrt()):.4f}")

    # DDPM tables + one posterior step per probe index, from the reference's DDPM class
    d = {}
    for Ns in (1000, 20):
        m = ref.sde_diffusion.DDPM(Ns)
        for k, v in m.state_dict().items():
            d[f"Ns{Ns}.{k}"] = v.numpy()
    m = ref.sde_diffusion.DDPM(1000)
    rs = np.random.RandomState(7)
    x = torch.from_numpy(rs.standard_normal((4, 3, 8, 8)).astype(np.float32))
    e = torch.from_numpy(rs.standard_normal((4, 3, 8, 8)).astype(np.float32))
    z = torch.from_numpy(rs.standard_normal((4, 3, 8, 8)).astype(np.float32))
    d["step.x"], d["step.eps"], d["step.z"] = x.numpy(), e.numpy(), z.numpy()
    for i in (0, 1, 500, 999):
        it = torch.full((4,), i, dtype=torch.long)
        x0 = torch.clip(m.predict_start_from_noise(x, it, e), -1, 1)
        mean, _, lv, _ = m.p_mean_variance(x0, x, it)
        d[f"step.out{i}"] = (mean + (0.5 * lv).exp() * z).numpy()
        d[f"step.t{i}"] = (1.0 * it / m.Ns).numpy()
    np.savez_compressed(os.path.join(HERE, "ddpm_reference.npz"), **d)
    print("ddpm tables written")


if __name__ == "__main__":
    main()
