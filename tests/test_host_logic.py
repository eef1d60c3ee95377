"""Host-side mirror of the reference interface (CPU): parameter naming, config defaults, time grids,
DDPM buffers, sharding arithmetic, world_size-2 gloo gather."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from golden_configs import GOLDEN_CONFIGS
from oracle import ddpm as D
from oracle import integrators as I
from oracle import unet as O


def _cfg_kwargs(cfg):
    return dict(image_size=cfg.image_size, in_channels=cfg.in_channels, model_channels=cfg.model_channels,
                out_channels=cfg.out_channels, num_res_blocks=cfg.num_res_blocks, attention_resolutions=cfg.attention_ds,
                channel_mult=cfg.channel_mult, num_heads=cfg.num_heads, num_head_channels=cfg.num_head_channels,
                num_heads_upsample=cfg.num_heads_upsample, use_scale_shift_norm=cfg.use_scale_shift_norm,
                resblock_updown=cfg.resblock_updown, use_new_attention_order=cfg.use_new_attention_order)


@pytest.mark.parametrize("name", list(GOLDEN_CONFIGS))
def test_state_dict_names_match_reference_layout(pkg, name):
    cfg = GOLDEN_CONFIGS[name][0]
    m = pkg.UNetModel(**_cfg_kwargs(cfg))
    shapes = O.param_shapes(cfg)       # verified against the vendored module in test_oracle.py
    sd = m.state_dict()
    assert list(sd.keys()) == list(shapes.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == shapes[k], k
    m.load_state_dict(O.seeded_params(cfg, 0))


def test_default_init_zero_modules(pkg):
    m = pkg.UNetModelWrapper(dim=(1, 28, 28), num_channels=32, num_res_blocks=1)
    sd = m.state_dict()
    assert float(sd["out.2.weight"].abs().max()) == 0.0
    assert float(sd["middle_block.1.proj_out.weight"].abs().max()) == 0.0
    assert float(sd["input_blocks.1.0.out_layers.3.weight"].abs().max()) == 0.0
    assert m.config.channel_mult == (1, 2, 2) and m.config.attention_ds == (1,)


def test_wrapper_constructors(pkg):
    m = pkg.UNetModelWrapper(dim=(3, 32, 32), num_res_blocks=2, num_channels=128, channel_mult=[1, 2, 2, 2], num_heads=4,
                             num_head_channels=64, attention_resolutions="16", dropout=0.1)
    assert sum(p.numel() for p in m.parameters()) == 35_746_307
    assert m.config.attention_ds == (2,)
    c = pkg.UNetModelWrapper(dim=(1, 28, 28), num_channels=32, num_res_blocks=1, num_classes=10, class_cond=True)
    assert sum(p.numel() for p in c.parameters()) == 1_076_641 and c.num_classes == 10
    i = pkg.InPaintModelWrapper(dim=(1, 28, 28), num_channels=32, num_res_blocks=1, num_classes=None, class_cond=True)
    assert i.config.in_channels == 2 and i.config.out_channels == 1 and i.num_classes is None
    s = pkg.SuperResModelWrapper(dim=(3, 64, 64), num_channels=128, num_res_blocks=1, num_classes=None, class_cond=True)
    assert s.config.in_channels == 6 and s.config.channel_mult == (1, 2, 3, 4)
    d = pkg.create_model(image_size=64, in_channels=3, out_channels=3, num_channels=128, num_res_blocks=1,
                         resblock_updown=True, num_head_channels=64, use_scale_shift_norm=True, num_heads=4)
    assert sum(p.numel() for p in d.parameters()) == 68_156_163
    with pytest.raises(ValueError):
        pkg.create_model(image_size=28, in_channels=1, out_channels=1, num_channels=32, num_res_blocks=1)


def test_checkpoint_ingest_formats(pkg, tmp_path):
    cfg = GOLDEN_CONFIGS["tiny_neworder"][0]
    p = O.seeded_params(cfg, 3)
    m = pkg.UNetModel(**_cfg_kwargs(cfg))
    pkg.load_checkpoint(m, {"ema_model": {f"module.{k}": v for k, v in p.items()}}, strict=True)   # compute_fid.py:54-64
    assert torch.equal(m.state_dict()["out.2.weight"], p["out.2.weight"])
    m2 = pkg.UNetModel(**_cfg_kwargs(cfg))
    path = str(tmp_path / "ck.pth")
    torch.save({"ema": {f"ema_model.{k}": v for k, v in p.items()} | {"step": torch.tensor(3)}}, path)   # unet.py:107-115
    pkg.load_checkpoint(m2, path)
    assert torch.equal(m2.state_dict()["time_embed.0.weight"], p["time_embed.0.weight"])


def test_euler_time_grid_matches_oracle(pkg):
    for n in (101, 100, 3, 1000):
        ts = torch.linspace(0, 1, n)
        assert pkg.euler_time_grid(ts) == I.euler_time_grid(ts)


def test_ddpm_buffers_match_oracle(pkg):
    for Ns in (1000, 20):
        m = pkg.DDPM(Ns)
        tb = D.ddpm_tables(Ns)
        for k, v in m.state_dict().items():
            assert torch.equal(v, tb[k]), k
        mt = m.model_time()
        assert all(float(mt[i]) == float(D.eps_time(i, Ns)) for i in (0, 1, Ns // 2, Ns - 1))


def test_unsupported_paths_raise(pkg):
    with pytest.raises(NotImplementedError):
        pkg.ReconstructionGuidance(gamma=10.0, start_fraction=1.0, update_rule="before", n_corrector=0, delta=0.1)
    with pytest.raises(TypeError):      # Replacement needs a likelihood that marks the holes (pad_value)
        pkg.get_conditional_sample_fn(lambda x, i: x, pkg.DDPM(10), pkg.Replacement(), pkg.HyperResolution(7, 7))
    with pytest.raises(NotImplementedError):
        pkg.NeuralODE(lambda t, x: x, solver="rk4")


def test_shard_range_partitions_exactly(pkg):
    for total in (0, 1, 7, 1024, 50000):
        for world in (1, 2, 3, 8):
            spans = [pkg.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        pkg.shard_range(10, 2, 2)


def _gloo_worker(rank, world, port, total, q):
    import __graft_entry__ as g
    pkg = g.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = pkg.shard_range(total, rank, world)
    full = (torch.arange(total * 3 * 2 * 2) % 251).to(torch.uint8).reshape(total, 3, 2, 2)
    got = pkg.gather_uint8(full[lo:hi].clone(), total)
    ok = torch.equal(got, full)
    # FID-style sufficient statistics: all-reduce of per-rank sums equals the global sum
    feats = full[lo:hi].double().reshape(hi - lo, -1)
    s = feats.sum(0)
    dist.all_reduce(s)
    ok = ok and torch.allclose(s, full.double().reshape(total, -1).sum(0))
    # the scalar all-reduce behind the shared dopri5 step controller (integrators._allreduce_sum)
    tot = pkg.integrators._allreduce_sum([float(rank + 1), 0.5], dist.group.WORLD)
    ok = ok and tot == [float(sum(range(1, world + 1))), 0.5 * world]
    q.put((rank, ok))
    dist.destroy_process_group()


def test_world_size_2_gloo_gather_ragged(pkg):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the engine arm) on a one-sample, one-NFE sample:
    one JSON line with the contract's keys, `impl: reference`, zero copy bytes and a cpu_baseline describing the run."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-batch", "1", "--ref-nfe", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"]


def test_vectorised_box_draw_consumes_the_generator_like_the_reference_loop(pkg):
    """InPainting.sample_boxes draws all (h, w) pairs with one torch.randint(lo, hi, (B, 2)); the reference draws h then
    w per sample with scalar calls (likelihoods.py:49-53, 78-87).  Same global CPU generator, same values, same order,
    and the same generator state afterwards."""
    lk = pkg.InPainting(patch_size=20, pad_value=-2.0)
    for B in (1, 7, 4096):
        torch.manual_seed(123)
        loop = torch.tensor([[int(torch.randint(5, 64 - 20 - 5, size=())), int(torch.randint(5, 64 - 20 - 5, size=()))]
                             for _ in range(B)], dtype=torch.int32)
        after_loop = torch.rand(1)
        torch.manual_seed(123)
        vec = lk.sample_boxes(B, 64)
        after_vec = torch.rand(1)
        assert vec.dtype == torch.int32 and tuple(vec.shape) == (B, 2)
        assert torch.equal(vec, loop) and torch.equal(after_loop, after_vec)
    assert tuple(lk.sample_boxes(0, 64).shape) == (0, 2)
    with pytest.raises(RuntimeError):      # 28 px with the default patch has an empty range
        pkg.InPainting(patch_size=20, pad_value=-2.0).sample_boxes(2, 28)


def test_plain_callable_eps_models_are_accepted(pkg):
    """The sampler factories take any callable (sde_diffusion.py:11 Network); construction needs no GPU."""
    ddpm = pkg.DDPM(10)
    f = pkg.get_prior_sample_fn(lambda xi, i: xi, ddpm, None, None)
    g2 = pkg.get_conditional_sample_fn(lambda xi, i: xi, ddpm, pkg.Replacement(), pkg.InPainting(4, -2.0))
    assert callable(f) and callable(g2)
    with pytest.raises(TypeError):
        pkg.get_prior_sample_fn(3, ddpm, None, None)
