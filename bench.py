#!/usr/bin/env python
"""Headline benchmark: CIFAR-10 CFM U-Net, 100-step Euler sampling, batch 1024 per GPU.

    python bench.py --gpus N --steps K --warmup W            # this repo's engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A "step" is one pass of the hot path over one batch: integrate `batch` synthetic noise images
through `nfe` Euler steps of the CIFAR U-Net (cifar10/compute_fid.py:73-88) and produce the uint8
images.  Prints ONE JSON line (contract in the task statement).  Multi-GPU: one process per GPU
(torchrun), the batch is sharded with no collective inside the loop (weak scaling: every rank runs
a full `batch`); NCCL only gathers the finished uint8 images in the end-to-end leg.
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
    model = pkg.UNetModelWrapper(dim=(3, 32, 32), num_res_blocks=2, num_channels=128, channel_mult=[1, 2, 2, 2], num_heads=4,
                                 num_head_channels=64, attention_resolutions="16", dropout=0.1, precision=args.precision)
    model.load_state_dict(O.seeded_params(cfg, 0))
    model = model.to(dev).eval()
    eng = model.engine()
    B, nfe = args.batch, args.nfe
    t_span = torch.linspace(0, 1, nfe + 1)
    use_graph = not args.no_graph
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident leg: `value` ----------------
    x_dev = [torch.randn(B, 3, 32, 32, device=dev, generator=gen) for _ in range(2)]

    def step_resident(i):
        return pkg.sample_euler(model, x_dev[i % 2], t_span, return_uint8=True, use_graph=use_graph)

    for i in range(args.warmup):
        step_resident(i)
    launches_per_step = eng.last_launches
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        ev0.record()
        for i in range(args.steps):
            step_resident(i)
        ev1.record()
        barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    value = B * world * args.steps / (ms_total / 1e3)

    # ---------------- end-to-end leg: host buffers in, uint8 images out ----------------
    x_host = [torch.randn(B, 3, 32, 32).pin_memory() for _ in range(2)]
    img_host = torch.empty(B, 3, 32, 32, dtype=torch.uint8).pin_memory()

    def step_e2e(i):
        xd = x_host[i % 2].to(dev, non_blocking=True)
        _, img = pkg.sample_euler(model, xd, t_span, return_uint8=True, use_graph=use_graph)
        img_host.copy_(img, non_blocking=True)

    for i in range(min(args.warmup, 1)):
        step_e2e(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_e2e(i)
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = B * world * args.steps / (float(ms2) / 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (conv_tc_kernel), measured live ----------------
    peaks = measured_peaks()
    rows = eng.profile_forward(x_dev[0], 0.5, repeats=3)
    tc = [r for r in rows if r["kind"] == "conv_tcgen05"]
    roofline = None
    kernel_ms = {}
    for r in rows:
        kernel_ms[r["kind"]] = kernel_ms.get(r["kind"], 0.0) + r["ms"]
    if tc:
        tc_ms = sum(r["ms"] for r in tc)
        tc_fl = sum(r["flops"] for r in tc)
        achieved = tc_fl / (tc_ms * 1e-3) / 1e12
        traffic = None
        tr_path = os.path.join(ROOT, "profiles", "conv_tc_traffic.json")
        if os.path.isfile(tr_path):
            try:
                traffic = json.load(open(tr_path)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roofline = {"bound": "tensor", "kernel": "conv_tc2_kernel / conv_tc_kernel (tcgen05 implicit-GEMM conv, all 64 launches of an NFE)", "achieved": achieved,
                    "peak": peaks["sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["sustained"], "traffic": traffic,
                    "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                    "launches_per_nfe": len(tc), "flops_per_launch_avg": tc_fl / len(tc), "ms_per_launch_avg": tc_ms / len(tc),
                    "share_of_nfe": tc_ms / sum(kernel_ms.values())}
    flops_step = eng.flops_per_sample * B * nfe
    whole = {"achieved_tflops": value / world * nfe * eng.flops_per_sample / 1e12,
             "frac_of_burst": value / world * nfe * eng.flops_per_sample / 1e12 / peaks["burst"],
             "frac_of_sustained": value / world * nfe * eng.flops_per_sample / 1e12 / peaks["sustained"],
             "ms_per_nfe": ms_total / args.steps / nfe, "kernel_ms_per_nfe": kernel_ms}

    cpu = None
    if world == 1 and not args.skip_cpu_baseline:
        v, sec, cores = time_cpu_oracle(8, 2, 2, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "oracle port (torch CPU fp32), batch 8 x 2 Euler NFE per step, 1 warm-up + 2 timed, extrapolated to 100 NFE"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": {"workload": "CIFAR-10 CFM UNetModel (num_channels=128, 2 res blocks, mult 1-2-2-2, attn@16x16, 4x64 heads), "
                               "100-step Euler, synthetic batch 1024 per GPU, random-init (seeded) weights",
                   "per_gpu_batch": B, "global_batch": B * world, "nfe": nfe, "parallelism": f"dp{world} (batch sharded, no collective in loop)",
                   "cuda_graph": use_graph,
                   "l2": f"activation working set {eng.workspace_bytes(B) / 2**30:.2f} GiB per NFE >> 126 MB L2 (no flush needed)"},
        "nfe_per_sec": value * nfe / B, "sample_nfe_per_sec": value * nfe,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * 3 * 32 * 32 * 4 * world,
                "d2h_bytes_per_step": B * 3 * 32 * 32 * world},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks.summary(),
        "roofline": roofline, "whole_step": whole, "flops_per_step": flops_step,
        "tensor_core_convs_per_nfe": eng.tensor_core_convs,
    }
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
