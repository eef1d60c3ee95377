#!/usr/bin/env python
"""Headline benchmark: CIFAR-10 CFM U-Net, 100-step Euler sampling, batch 1024 per GPU.

    python bench.py --gpus N --steps K --warmup W            # this repo's engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A "step" is one pass of the hot path over one batch: integrate `batch` synthetic noise images
through `nfe` Euler steps of the CIFAR U-Net (cifar10/compute_fid.py:73-88) and produce the uint8
images.  Prints ONE JSON line (contract in the task statement).  Multi-GPU: one process per GPU
(torchrun), the batch is sharded with no collective inside the loop.  N > 1 runs two legs: weak
scaling (every rank integrates a full `batch`; the headline `value` unless --scaling strong) and
strong scaling (`batch` samples in total, batch/N per rank; reported under "strong_scaling").
In the end-to-end leg of N > 1 the finished uint8 shards are all-gathered over NCCL
(`distributed.gather_uint8`) and rank 0 copies the whole set to the host inside the timed region.
After the timed region rank 0 checks one NFE at the benchmarked batch against the reference's
golden vector ("parity_check") and reports the final-sample drift of the 100-step sampler
(bf16 engine vs fp32 engine vs CPU oracle, "drift").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "CIFAR-10 CFM samples/sec (100 Euler NFE, bs 1024)"
UNIT = "samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="samples per GPU per step")
    ap.add_argument("--nfe", type=int, default=100, help="Euler steps (integration_steps)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--ref-batch", type=int, default=16)
    ap.add_argument("--ref-nfe", type=int, default=2)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: which leg is the headline `value` (weak: --batch samples per GPU; strong: --batch samples in "
                         "total, batch/N per GPU).  Both are measured and printed either way.")
    ap.add_argument("--skip-drift", action="store_true")
    ap.add_argument("--drift-batch", type=int, default=16, help="samples of the bf16-vs-fp32 100-step drift report")
    ap.add_argument("--drift-oracle-batch", type=int, default=4, help="of which this many also run through the CPU oracle")
    return ap.parse_args()


def cifar_config():
    from oracle import unet as O
    return O.config_from_wrapper((3, 32, 32), 128, 2, channel_mult=[1, 2, 2, 2], num_heads=4, num_head_channels=64,
                                 attention_resolutions="16")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            d = json.load(open(path))
            return {"burst": float(d["bf16_tflops"]), "sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    "hbm": float(d["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
        except Exception:
            pass
    return {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
# CPU baseline = the oracle (a port of the reference's PyTorch path), on the host cores
# ------------------------------------------------------------------------------------------------
def time_cpu_oracle(ref_batch: int, ref_nfe: int, steps: int, warmup: int):
    """Returns (samples/s extrapolated to 100 NFE, seconds per step, cores)."""
    import torch
    from oracle import integrators as I
    from oracle import unet as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = cifar_config()
    params = O.seeded_params(cfg, 0)
    g = torch.Generator().manual_seed(0)
    f = lambda t, x: O.wrapper_forward(cfg, params, t, x)
    t_span = torch.linspace(0, 1, 101)[: ref_nfe + 1]          # the first ref_nfe steps of the 100-step grid
    times = []
    for it in range(warmup + steps):
        x0 = torch.randn(ref_batch, 3, 32, 32, generator=g)
        t0 = time.perf_counter()
        traj = I.euler_trajectory(f, x0, t_span)
        _ = (traj[-1] * 127.5 + 128).clip(0, 255).to(torch.uint8)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec_per_step = sum(times) / len(times)
    sec_per_sample_nfe = sec_per_step / (ref_batch * ref_nfe)
    return 1.0 / (sec_per_sample_nfe * 100.0), sec_per_step, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, sec, cores = time_cpu_oracle(args.ref_batch, args.ref_nfe, args.steps, max(args.warmup, 1))
    sample = (f"oracle port (reference U-Net arithmetic, torch CPU fp32), batch {args.ref_batch} x {args.ref_nfe} Euler NFE "
              f"per step, extrapolated linearly to 100 NFE")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "CIFAR-10 CFM UNetModel ch128 x2 res blocks, attn@16x16, 100-step Euler, batch 1024 "
                               "(CPU arm: bounded sample, see cpu_baseline.sample)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.stop = gpu_index, [], threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([v.strip() for v in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th.start(); return self

    def __exit__(self, *a):
        self.stop.set(); self.th.join(timeout=6)

    def summary(self):
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        pw = [float(r[3]) for r in self.rows if len(r) > 3 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows)}


def _drift_report(pkg, O, model_bf16, cfg, params, dev, nfe, batch, oracle_batch):
    """Final-sample drift of the 100-step Euler sampler (north_star: "final-sample drift must be reported"):
    the bf16 engine against the fp32 engine on `batch` samples and both against the CPU oracle on the first
    `oracle_batch` of them (cifar10/compute_fid.py:73-88: same x0, same t_span, uint8 conversion included)."""
    import torch
    from oracle import ddpm as D
    from oracle import integrators as I
    t_span = torch.linspace(0, 1, nfe + 1)
    x0 = torch.randn(batch, 3, 32, 32, generator=torch.Generator().manual_seed(2024))
    m32 = pkg.UNetModelWrapper(dim=(3, 32, 32), num_res_blocks=2, num_channels=128, channel_mult=[1, 2, 2, 2], num_heads=4,
                               num_head_channels=64, attention_resolutions="16", dropout=0.1, precision="fp32")
    m32.load_state_dict(params)
    m32 = m32.to(dev).eval()
    xb, ib = pkg.sample_euler(model_bf16, x0.to(dev), t_span, return_uint8=True, use_graph=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    xf, if_ = pkg.sample_euler(m32, x0.to(dev), t_span, return_uint8=True, use_graph=False)
    torch.cuda.synchronize()
    fp32_s = time.perf_counter() - t0
    m32.refresh()
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    u8 = lambda a, b: {"mismatch_frac": float((a != b).float().mean()), "max_abs_diff": int((a.int() - b.int()).abs().max())}
    rep = {"nfe": nfe, "batch": batch, "bf16_vs_fp32_engine": {"rel_l2": rel(xb.cpu(), xf.cpu()), "uint8": u8(ib.cpu(), if_.cpu())},
           # throughput of the exact mode (fp32 storage, CUDA-core kernels, <= 1e-4 per NFE): this very run, wall clock
           "fp32_exact_mode": {"samples_per_s": batch / fp32_s, "batch": batch, "nfe": nfe,
                               "note": "correctness mode on CUDA cores (no tensor-core fp32 path); first call includes engine build"}}
    if oracle_batch > 0:
        t0 = time.perf_counter()
        xo = I.euler_trajectory(lambda t, x: O.wrapper_forward(cfg, params, t, x), x0[:oracle_batch], t_span)[-1]
        io = D.to_uint8(xo)
        rep["oracle_batch"] = oracle_batch
        rep["oracle_seconds"] = time.perf_counter() - t0
        rep["bf16_vs_oracle"] = {"rel_l2": rel(xb[:oracle_batch].cpu(), xo), "uint8": u8(ib[:oracle_batch].cpu(), io)}
        rep["fp32_engine_vs_oracle"] = {"rel_l2": rel(xf[:oracle_batch].cpu(), xo), "uint8": u8(if_[:oracle_batch].cpu(), io)}
    return rep


def _parity_check(model, dev, B):
    """One NFE at the benchmarked batch with the reference's golden rows planted at both ends: bit-equality with the
    batch-2 evaluation and rel-L2 against the reference's own output (tests/golden/unet_cifar.npz)."""
    import numpy as np
    import torch
    g = np.load(os.path.join(ROOT, "tests", "golden", "unet_cifar.npz"))
    gx, gt = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["t"]).to(dev)
    eng = model.engine()
    small = eng.forward(gx, gt)
    gen = torch.Generator(device=dev).manual_seed(7)
    x = torch.randn(B, 3, 32, 32, device=dev, generator=gen)
    t = torch.rand(B, device=dev, generator=gen)
    spots = sorted({0, max(B - 2, 0)})
    for s_ in spots:
        x[s_:s_ + 2] = gx; t[s_:s_ + 2] = gt
    out = eng.forward(x, t)
    want = torch.from_numpy(g["out"])
    rel = float((out[:2].cpu().double() - want.double()).norm() / want.double().norm())
    return {"batch": B, "rel_l2_vs_reference_golden": rel, "tol": 2e-2, "within_tol": rel < 2e-2,
            "bit_equal_vs_batch2": bool(all(torch.equal(out[s_:s_ + 2], small) for s_ in spots)),
            "finite": bool(torch.isfinite(out).all())}


def run_engine(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    from oracle import unet as O

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl engine needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    pkg = g.load_package()

    cfg = cifar_config()
    params = O.seeded_params(cfg, 0)
    model = pkg.UNetModelWrapper(dim=(3, 32, 32), num_res_blocks=2, num_channels=128, channel_mult=[1, 2, 2, 2], num_heads=4,
                                 num_head_channels=64, attention_resolutions="16", dropout=0.1, precision=args.precision)
    model.load_state_dict(params)
    model = model.to(dev).eval()
    eng = model.engine()
    nfe = args.nfe
    t_span = torch.linspace(0, 1, nfe + 1)
    use_graph = not args.no_graph
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        v = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v)

    def run_legs(Bper, total, steps, warmup, clocks_wanted):
        """Device-resident leg and end-to-end leg at `Bper` samples on this rank (`total` over all ranks).
        Returns (ms_resident, ms_e2e, launches_per_step, clock summary | None), times = max over ranks."""
        x_dev = [torch.randn(Bper, 3, 32, 32, device=dev, generator=gen) for _ in range(2)]
        for i in range(warmup):
            pkg.sample_euler(model, x_dev[i % 2], t_span, return_uint8=True, use_graph=use_graph)
        launches = eng.last_launches
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clocks = ClockSampler(local) if clocks_wanted else None
        if clocks:
            clocks.__enter__()
        ev0.record()
        for i in range(steps):
            pkg.sample_euler(model, x_dev[i % 2], t_span, return_uint8=True, use_graph=use_graph)
        ev1.record()
        barrier()
        if clocks:
            clocks.__exit__()
        ms_res = max_over_ranks(ev0.elapsed_time(ev1))

        # end to end: pinned host noise in, uint8 images of ALL ranks out on rank 0's host (N > 1: NCCL all-gather of the
        # finished uint8 shards, cifar10/compute_fid.py:92-100 consumes the images on the host)
        x_host = [torch.randn(Bper, 3, 32, 32).pin_memory() for _ in range(2)]
        img_host = torch.empty(total if rank == 0 else 0, 3, 32, 32, dtype=torch.uint8).pin_memory()

        def step_e2e(i):
            xd = x_host[i % 2].to(dev, non_blocking=True)
            _, img = pkg.sample_euler(model, xd, t_span, return_uint8=True, use_graph=use_graph)
            if world > 1:
                img = pkg.gather_uint8(img, total)
            if rank == 0:
                img_host.copy_(img, non_blocking=True)

        for i in range(min(warmup, 1)):
            step_e2e(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step_e2e(i)
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        return ms_res, ms_e2e, launches, (clocks.summary() if clocks else None)

    B = args.batch
    legs = {}
    modes = ["weak"] if world == 1 else ["weak", "strong"]
    if world > 1 and args.scaling == "strong":
        modes = ["strong", "weak"]
    for mode in modes:
        if mode == "weak":
            Bper, total = B, B * world
        else:
            lo, hi = pkg.shard_range(B, rank, world)
            Bper, total = hi - lo, B
        ms_res, ms_e2e, launches, clk = run_legs(Bper, total, args.steps, args.warmup, mode == modes[0])
        legs[mode] = {"per_gpu_batch": Bper, "global_batch": total, "ms_per_step": ms_res / args.steps,
                      "value": total * args.steps / (ms_res / 1e3), "e2e_value": total * args.steps / (ms_e2e / 1e3),
                      "launches_per_step": launches, "clocks": clk,
                      "h2d_bytes_per_step": total * 3 * 32 * 32 * 4, "d2h_bytes_per_step": total * 3 * 32 * 32}
    head = legs[modes[0]]
    Bper = head["per_gpu_batch"]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (tcgen05 conv), measured live ----------------
    peaks = measured_peaks()
    x_prof = torch.randn(Bper, 3, 32, 32, device=dev, generator=gen)
    rows = eng.profile_forward(x_prof, 0.5, repeats=3)
    tc = [r for r in rows if r["kind"] == "conv_tcgen05"]
    roofline = None
    kernel_ms = {}
    for r in rows:
        kernel_ms[r["kind"]] = kernel_ms.get(r["kind"], 0.0) + r["ms"]
    if tc:
        tc_ms = sum(r["ms"] for r in tc)
        tc_fl = sum(r["flops"] for r in tc)
        tc_ex = sum(r["flops_executed"] for r in tc)
        achieved = tc_fl / (tc_ms * 1e-3) / 1e12
        executed = tc_ex / (tc_ms * 1e-3) / 1e12
        traffic, traffic_src = None, None
        for name in ("r02_conv_tc_traffic.json", "conv_tc_traffic.json"):
            tr_path = os.path.join(ROOT, "profiles", name)
            if os.path.isfile(tr_path):
                try:
                    traffic = json.load(open(tr_path)).get("dram_bytes_per_launch")
                    traffic_src = f"profiles/{name} (ncu --set full capture of an earlier run of this command, not this run)"
                    break
                except Exception:
                    traffic = None
        roofline = {"bound": "tensor", "kernel": "conv_tc2_kernel / conv_tc_kernel (tcgen05 implicit-GEMM conv, all launches of an NFE)",
                    "achieved": achieved, "peak": peaks["burst"], "unit": "TFLOP/s", "frac": achieved / peaks["burst"],
                    "achieved_algorithmic": achieved, "achieved_executed": executed,
                    "frac_executed": executed / peaks["burst"],
                    "frac_of_sustained": achieved / peaks["sustained"], "frac_executed_of_sustained": executed / peaks["sustained"],
                    "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": peaks["source"] + ", burst figure (per-launch CUDA-event timings, kernels timed alone)",
                    "note": "algorithmic = 2*MAC of the reference's convs; executed = 2*MAC the kernels issue "
                            "(the folded nearest-x2-upsample convs issue 4/9, the stem/head GEMMs run zero-padded K/N)",
                    "launches_per_nfe": len(tc), "flops_per_launch_avg": tc_fl / len(tc), "ms_per_launch_avg": tc_ms / len(tc),
                    "share_of_nfe": tc_ms / sum(kernel_ms.values())}
    gn = [r for r in rows if r["kind"] == "groupnorm"]
    hbm = None
    if gn:
        gn_ms = sum(r["ms"] for r in gn)
        gn_b = sum(r["bytes"] for r in gn)
        hbm = {"kernel": "GroupNorm(+FiLM)(+SiLU) pass, all launches of an NFE", "bound": "hbm", "achieved": gn_b / (gn_ms * 1e-3) / 1e9,
               "peak": peaks["hbm"], "unit": "GB/s", "frac": gn_b / (gn_ms * 1e-3) / 1e9 / peaks["hbm"],
               "launches_per_nfe": len(gn), "ms_per_nfe": gn_ms, "bytes_per_nfe": gn_b}
    nfe_tf = head["value"] / world * nfe * eng.flops_per_sample / 1e12
    whole = {"achieved_tflops": nfe_tf, "frac_of_burst": nfe_tf / peaks["burst"], "frac_of_sustained": nfe_tf / peaks["sustained"],
             "ms_per_nfe": head["ms_per_step"] / nfe, "kernel_ms_per_nfe": kernel_ms}

    parity = _parity_check(model, dev, Bper)
    drift = None
    cpu = None
    if world == 1:
        if not args.skip_drift:
            drift = _drift_report(pkg, O, model, cfg, params, dev, nfe, args.drift_batch, args.drift_oracle_batch)
        if not args.skip_cpu_baseline:
            v, sec, cores = time_cpu_oracle(8, 2, 2, 1)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "oracle port (torch CPU fp32), batch 8 x 2 Euler NFE per step, 1 warm-up + 2 timed, extrapolated to 100 NFE"}

    line = {
        "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": modes[0], "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": {"workload": "CIFAR-10 CFM UNetModel (num_channels=128, 2 res blocks, mult 1-2-2-2, attn@16x16, 4x64 heads), "
                               "100-step Euler, synthetic batch 1024 per GPU, random-init (seeded) weights",
                   "per_gpu_batch": Bper, "global_batch": head["global_batch"], "nfe": nfe,
                   "parallelism": f"dp{world} (batch sharded, no collective in loop; NCCL all-gather of the uint8 images after it in the e2e leg)",
                   "cuda_graph": use_graph,
                   "l2": f"activation working set {eng.workspace_bytes(Bper) / 2**30:.2f} GiB per NFE >> 126 MB L2 (no flush needed)"},
        "nfe_per_sec": head["value"] * nfe / head["global_batch"] * world, "sample_nfe_per_sec": head["value"] * nfe,
        "e2e": {"value": head["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": head["h2d_bytes_per_step"],
                "d2h_bytes_per_step": head["d2h_bytes_per_step"]},
        "gpu_launches": head["launches_per_step"] * args.steps,
        "clocks": head["clocks"],
        "roofline": roofline, "hbm_roofline": hbm, "whole_step": whole, "flops_per_step": eng.flops_per_sample * Bper * nfe,
        "tensor_core_convs_per_nfe": eng.tensor_core_convs,
        "parity_check": parity,
    }
    for mode in modes[1:]:
        o = legs[mode]
        line[f"{mode}_scaling"] = {"value": o["value"], "unit": UNIT, "per_gpu_batch": o["per_gpu_batch"], "global_batch": o["global_batch"],
                                   "ms_per_step": o["ms_per_step"], "e2e_value": o["e2e_value"]}
    if drift is not None:
        line["drift"] = drift
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
