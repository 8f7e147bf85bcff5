#!/usr/bin/env python3
"""bench.py -- headline metric of BASELINE.json: full-ALD channel-estimates/sec on B200.

Workload (config[1], "CDL-C Fig-5c curve, batch=256, full sigma schedule, 1xB200"): one *step* is one
pass of the hot path over one batch of 256 synthetic CDL-shaped 16x64 channels (Np = 38 pilots, SNR
points of the Fig-5c sweep spread over the batch), every one taken through the complete schedule of
2311 sigma levels x 3 Langevin steps = 6933 fused network evaluations + updates, random-init ngf=8
weights of the shipped architecture (data: synthetic).  Under torchrun each rank runs its own batch of
256 (weak scaling, no data-path collective) and the per-rank NMSE logs are all-gathered over NCCL once
at the end of the step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--levels L]

`--levels` (default: full 2311) exists for quick functional runs only; the reported line always states
the schedule it ran.  `--impl reference` times the CPU restatement of the reference path (the oracle
port; the reference itself is Python-on-torch-CUDA and cannot travel to the GPU box) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

SIGMA_END = 2.599515446446343e-4
NUM_LEVELS, STEPS_EACH = 2311, 3
NT, NR, NP = 64, 16, 38
SNR_RANGE = np.arange(-10, 32.5, 2.5)           # test_score.py:72
ALPHA_STEP, BETA = 3e-11, 0.01                  # test_score.py:46-48
METRIC = "channel-estimates/sec (full ALD, 16x64 CDL-C)"


def make_batch(B, rank=0):
    from score_based_channels_b200 import synth
    H = synth.cdl_like_channels(B, NT, NR, seed=4321 + 100000 * rank)
    P = synth.qpsk_pilots(B, NT, NP, seed=1234 + rank)
    snr = SNR_RANGE[np.arange(B) % len(SNR_RANGE)]
    nv = synth.snr_to_noise_var(snr, NT).astype(np.float32)
    Y = synth.received_pilots(P, H, nv, seed=99 + rank)
    X0 = synth.cn01((B, NT, NR), np.random.default_rng(7 + rank))
    return P, Y, X0, H, nv


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md, clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10)
                self.rows.append([c.strip() for c in out.stdout.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=15)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def cpu_arm(args, levels, sample_b=None, max_seconds=25.0):
    """The reference path restated on the CPU (oracle port), all host threads, bounded sample."""
    from oracle import oracle as orc
    from score_based_channels_b200 import params
    sd = params.random_state(8, seed=1)
    cores = os.cpu_count() or 1
    orc.set_num_threads(cores)
    B = sample_b or max(cores, 8)
    P, Y, X0, H, nv = make_batch(B)
    net = orc.OracleNet(sd, 8, NT, NR)
    kw = dict(noise_var=nv, alpha_step=ALPHA_STEP, beta=BETA, sigma_end=SIGMA_END, steps_each=STEPS_EACH, seed=1)
    t0 = time.perf_counter()
    net.ald(P, Y, X0, H, level_begin=0, level_end=1, **kw)          # warm-up + calibration: 3 steps
    t_cal = (time.perf_counter() - t0) / STEPS_EACH
    n_lvl = int(max(1, min(levels, max_seconds / max(t_cal * STEPS_EACH, 1e-6))))
    t0 = time.perf_counter()
    net.ald(P, Y, X0, H, level_begin=0, level_end=n_lvl, **kw)
    dt = time.perf_counter() - t0
    t_step = dt / (n_lvl * STEPS_EACH)                              # seconds per Langevin step of the sample batch
    est_per_s = B / (t_step * levels * STEPS_EACH)
    return {"value": est_per_s, "unit": "estimates/s", "cores": cores, "kind": "port",
            "sample": "B=%d x %d levels x %d steps timed (%.1f s), per-step cost extrapolated to %d levels"
                      % (B, n_lvl, STEPS_EACH, dt, levels), "ms_per_langevin_step": t_step * 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--levels", type=int, default=NUM_LEVELS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the short run of the other precision mode")
    ap.add_argument("--precision", default=os.environ.get("SBC_PRECISION", "tf32x3"), choices=["tf32x3", "tf32"])
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    levels = args.levels
    config = {"workload": "Fig-5c CDL-C synthetic, batch=%d per GPU, %d sigma levels x %d steps, Nt=64 Nr=16 Np=38, "
                          "17 SNR points spread over the batch" % (args.batch, levels, STEPS_EACH),
              "global_batch": args.batch * world, "parallelism": "batch-sharded x%d" % world,
              "l2_policy": "per-step working set is re-streamed weights (1.48 MB) + per-sample state; inputs are "
                           "rewritten by the H2D copy every step in the e2e leg"}

    if args.impl == "reference":
        if rank != 0:
            return
        times = []
        for i in range(args.warmup + args.steps):
            r = cpu_arm(args, levels, max_seconds=12.0)
            if i >= args.warmup:
                times.append(r)
        v = float(np.mean([r["value"] for r in times]))
        cb = dict(times[-1]); cb["value"] = v
        line = {"metric": METRIC, "value": v, "unit": "estimates/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * args.batch / v, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "impl": "reference", "cpu_baseline": cb,
                "e2e": {"value": v, "unit": "estimates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    if local_rank == 0:
        ge.build()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    from score_based_channels_b200 import params, sampler
    from score_based_channels_b200.models import make_model

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    sd = params.random_state(8, seed=1)
    model = make_model(sd, ngf=8, precision=args.precision).to(dev)
    B = args.batch
    P, Y, X0, H, nv = make_batch(B, rank)
    host = [torch.from_numpy(a).pin_memory() for a in (P, Y, X0, H, nv)]
    ids = torch.arange(rank * B, (rank + 1) * B, dtype=torch.int64, device=dev)
    kw = dict(alpha_step=ALPHA_STEP, beta=BETA, sigma_end=SIGMA_END, level_begin=0, level_end=levels,
              steps_each=STEPS_EACH, seed=2026)
    nsteps_ald = levels * STEPS_EACH
    gather_buf = [torch.empty((nsteps_ald, B), dtype=torch.float32, device=dev) for _ in range(world)] if world > 1 else None

    def gpu_step(dP, dY, dX0, dH, dnv):
        X, nlog = sampler.ald_run(model, dP, dY, dX0, dH, noise_var=dnv, sample_ids=ids, **kw)
        if world > 1:      # the single collective of the path: gather the per-rank NMSE logs
            dist.all_gather(gather_buf, nlog)
        return X, nlog

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    pm = model.packed(NT, NR, dev)
    launches0 = pm.info().kernel_launches

    # ---- leg 1: inputs resident in HBM ("value") ----
    dres = [t.to(dev) for t in host]
    res_fn = lambda: gpu_step(*dres)
    for _ in range(args.warmup):
        res_fn()
    with ClockSampler(local_rank) as cs:
        ms_res = timed(res_fn, args.steps)
    clocks = cs.summary()
    launches = pm.info().kernel_launches - launches0 - args.warmup

    # ---- leg 2: end to end through the public API with HOST buffers ("e2e") ----
    out_host = [torch.empty((B, NT, NR), dtype=torch.complex64).pin_memory(),
                torch.empty((nsteps_ald, B), dtype=torch.float32).pin_memory()]

    def e2e_fn():
        d = [t.to(dev, non_blocking=True) for t in host]
        X, nlog = gpu_step(*d)
        out_host[0].copy_(X, non_blocking=True)
        out_host[1].copy_(nlog, non_blocking=True)

    for _ in range(max(1, args.warmup // 3)):
        e2e_fn()
    ms_e2e = timed(e2e_fn, args.steps)
    h2d = sum(t.numel() * t.element_size() for t in host)
    d2h = sum(t.numel() * t.element_size() for t in out_host)

    total_est = B * world * args.steps * (levels / NUM_LEVELS)      # full-ALD-equivalent estimates
    value = total_est / (ms_res * 1e-3)
    e2e = total_est / (ms_e2e * 1e-3)

    # ---- roofline of the dominant (only) kernel: sbc_ald_kernel, timed inside the long step ----
    peaks, how = measured_peaks()
    flops_per_launch = float(pm.prog.conv_flops) * nsteps_ald * B          # dense conv FLOP, reference convention
    kern_s = ms_res * 1e-3 / args.steps                                    # one launch per step dominates the step
    achieved = flops_per_launch / kern_s / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    # context for the fraction: this kernel's MMAs are mma.sync TF32, measured at one m16n8k8 per ~12 clk per SM
    # sub-partition (~683 FLOP/clk/SM, ~199 TFLOP/s per B200), and the fp32-equivalent mode
    # issues 3 of them per algorithmic MMA
    sm_clk_ghz = 1.965
    mma_sync_tf32_peak = (2 * 1024 * 4 / 12.0) * 148 * sm_clk_ghz / 1e3
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": 11.9e6 if levels == 16 else None,
                "traffic_note": "ncu dram__bytes_read+write = 11.9 MB for the 16-level launch at B=256 captured in "
                                "profiles/ (inputs once, log/outputs stay in L2): HBM is idle, the kernel lives in "
                                "shared memory; not captured for other schedule lengths",
                "peak_source": how + ", sustained dense bf16",
                "kernel": "sbc_ald_kernel<smem arena>, conv arithmetic = %s" % args.precision,
                "executed_mma_tflops": achieved * (3.0 if args.precision == "tf32x3" else 1.0),
                "mma_sync_tf32_ceiling_tflops": mma_sync_tf32_peak}

    # ---- the other precision mode, one short timed step (reported, not the headline) ----
    alt = None
    if world == 1 and not args.no_alt:
        ap_name = "tf32" if args.precision == "tf32x3" else "tf32x3"
        alt_model = make_model(sd, ngf=8, precision=ap_name).to(dev)
        alt_kw = dict(kw, level_end=min(levels, 96))
        alt_fn = lambda: sampler.ald_run(alt_model, *dres[:4], noise_var=dres[4], sample_ids=ids, **alt_kw)
        alt_fn()
        ms_alt = timed(alt_fn, 2)
        alt = {"precision": ap_name, "value": B * 2 * (alt_kw["level_end"] / NUM_LEVELS) / (ms_alt * 1e-3),
               "unit": "estimates/s", "sample": "2 steps of %d levels" % alt_kw["level_end"]}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_arm(args, levels)
        line = {"metric": METRIC, "value": value, "unit": "estimates/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": {"tf32x3": "tf32x3 (3xTF32 split, fp32-equivalent), f32 accumulate",
                          "tf32": "tf32 operands, f32 accumulate"}[args.precision],
                "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e, "unit": "estimates/s", "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / args.steps},
                "roofline": roofline, "cpu_baseline": cpu, "alt_precision": alt}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
