"""Two ranks over NCCL on two GPUs of one box: the path's only collectives - the all-gather of the finished uint8
images (cifar10/compute_fid.py:92-100) and the scalar all-reduce behind the shared dopri5 step controller -
run on the real backend (the gloo twins of these tests live in tests/test_host_logic.py).  Skipped on a one-GPU box."""
import os

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import __graft_entry__ as g
    from oracle import integrators as I
    from oracle import unet as O
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        pkg = g.load_package()
        ok = {}
        # ragged and even all-gathers of uint8 shards
        for total in (7, 8, 1024):
            full = (torch.arange(total * 3 * 4 * 4) % 251).to(torch.uint8).reshape(total, 3, 4, 4)
            lo, hi = pkg.shard_range(total, rank, world)
            got = pkg.gather_uint8(full[lo:hi].to(dev), total)
            ok[f"gather{total}"] = bool(torch.equal(got.cpu(), full))
        # sharded Euler sampling: every rank integrates its slice, images are gathered over NCCL
        cfg = O.config_from_wrapper((3, 16, 16), 32, 1, channel_mult=(1, 2), attention_resolutions="8", num_heads=2)
        params = O.seeded_params(cfg, 31)
        m = pkg.UNetModelWrapper(dim=(3, 16, 16), num_channels=32, num_res_blocks=1, channel_mult=(1, 2),
                                 attention_resolutions="8", num_heads=2, precision="fp32")
        m.load_state_dict(params)
        m = m.to(dev).eval()
        x0 = torch.randn(5, 3, 16, 16, generator=torch.Generator().manual_seed(3))
        t_span = torch.linspace(0, 1, 5)
        _, imgs = pkg.sample_euler_sharded(m, x0, t_span, use_graph=True)
        _, want = pkg.sample_euler(m, x0.to(dev), t_span, return_uint8=True, use_graph=False)
        ok["sharded_euler"] = bool(torch.equal(imgs, want))
        # dopri5 with one step controller across the ranks (NCCL branch of integrators._allreduce_sum)
        t = torch.tensor([0.0, 1.0])
        st_sh, st_one = {}, {}
        y_sh = pkg.odeint_sharded(m, x0, t, rtol=1e-4, atol=1e-4, stats=st_sh)[-1]
        y_one = pkg.odeint(m, x0.to(dev), t.to(dev), rtol=1e-4, atol=1e-4, method="dopri5", stats=st_one)[-1]
        lo, hi = pkg.shard_range(5, rank, world)
        ok["dopri5_steps"] = st_sh["steps"] == st_one["steps"] and st_sh["accepted"] == st_one["accepted"]
        ok["dopri5_state"] = bool((y_sh - y_one[lo:hi]).abs().max() < 1e-5)
        tot = pkg.integrators._allreduce_sum([float(rank + 1), 0.25], dist.group.WORLD)
        ok["allreduce"] = tot == [float(sum(range(1, world + 1))), 0.25 * world]
        q.put((rank, ok))
        dist.destroy_process_group()
    except Exception as exc:   # surface the failure in the parent instead of a queue timeout
        import traceback
        q.put((rank, {"exception": traceback.format_exc()}))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL, one rank per GPU)")
def test_two_ranks_nccl_gather_and_shared_controller(pkg):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(120)
    for rank in (0, 1):
        assert "exception" not in res[rank], res[rank]["exception"]
        assert all(res[rank].values()), (rank, res[rank])
