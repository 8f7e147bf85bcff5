#!/usr/bin/env python3
"""bench.py -- headline metric of BASELINE.json: full-ALD channel-estimates/sec on B200.

A *step* is one pass of the hot path (one fused launch of the annealed-Langevin kernel through the C ABI) over one batch
of synthetic CDL-shaped channels taken through the sigma schedule: per level 3 x (NCSNv2Deepest forward + data-consistency
gradient + Langevin update + Philox noise + NMSE), random-init ngf=8 weights of the shipped architecture (data: synthetic).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 2|3|4|5] [--engine auto|1|2]
                  [--batch B] [--levels L] [--no-extra]

Workloads (BASELINE.json `configs`):
  2 (default, the headline): "CDL-C Fig-5c curve, batch=256, full sigma schedule, 1xB200" -- 256 channels per GPU, 17 SNR
    points spread over the batch, 2311 levels x 3 steps.  Weak scaling: every rank runs its own 256.
  3: "SNR sweep -10..30 dB x 512 realisations, batch=4096" -- one launch of B=4096 per GPU with a per-sample noise_var.
  4: "tune_hparams_score alpha x beta grid" -- a FIXED total of 20 400 trajectories (4 alpha x 3 beta x 17 SNR x 100
    channels) sharded over the ranks: strong scaling.
  5: "Nt=128 Nr=32, batch=8192" -- 8192 trajectories per GPU at 128x32 (Np=76), weak scaling.
Configs 3-5 take minutes per full-schedule step; when they ride along with the default run (`extra`) they are timed on a
truncated schedule (`levels_run`), which the line states -- the per-level cost does not depend on the level.

Engines: 1 = fused shared-memory-arena kernel on mma.sync 3xTF32 (`tf32x3`); 2 = tcgen05 / TMEM kernel with fp16 hi/lo
split operands (`fp16x2`).  Both are fp32-equivalent (forward error ~2e-6 against the reference).  `auto` picks the
engine that is faster for the workload (measured: engine 1 at 64x16 where a sample fits one SM's shared memory, engine 2 at 128x32).
`--impl reference` times the CPU restatement of the reference path (the oracle port; the reference itself is Python on
torch-CUDA and cannot travel to the GPU box) on the host cores.
"""
import argparse
import ctypes as C
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
SNR_RANGE = np.arange(-10, 32.5, 2.5)           # test_score.py:72
ALPHA_STEP, BETA = 3e-11, 0.01                  # test_score.py:46-48
METRIC = "channel-estimates/sec (full ALD, 16x64 CDL-C)"
ENGINE_PRECISION = {1: "tf32x3", 2: "fp16x2"}
# HBM traffic of one launch, from the ncu captures committed under profiles/ (dram__bytes_read + write):
# bytes = fixed + per_level * levels at the captured batch; scaled linearly with the batch.
TRAFFIC = {1: {"batch": 256, "fixed": 15.0e6, "per_level": 0.378e6, "src": "profiles/r02b_ncu_engine1_2cta_raw.txt (16-level launch: 11.9 MB read = the "
                                                                         "inputs, 9.2 MB written) and a 2-level capture (4.5 MB written at B=296): "
                                                                         "the per-CTA park areas (41 MB in all) live in L2, a trickle of dirty lines is written back"},
           2: {"batch": 256, "fixed": 0.21e9, "per_level": 0.833e9, "src": "profiles/r02_ncu_engine2_raw.txt and r02_ncu_engine2_2levels_raw.txt "
                                                                         "(16 / 2 levels: 13.5 / 1.88 GB: the 115 MB working set of 256 per-CTA arenas "
                                                                         "does not stay in the 126 MB L2, dirty lines are written back)"}}


def workload(cfg, world, rank, batch=None):
    """(B_local, Nt, Nr, Np, description, scaling, global_batch)"""
    if cfg == 2:
        B = batch or 256
        return B, 64, 16, 38, "Fig-5c CDL-C synthetic, batch=%d per GPU, 17 SNR points spread over the batch" % B, "weak", B * world
    if cfg == 3:
        B = batch or 4096
        return B, 64, 16, 38, "SNR sweep -10..30 dB x 512 realisations in launches of batch=%d per GPU, per-sample noise_var" % B, "weak", B * world
    if cfg == 4:
        total = batch or 20400
        per = (total + world - 1) // world
        lo = min(rank * per, total)
        B = min(lo + per, total) - lo
        return B, 64, 16, 38, "tune_hparams_score grid: 4 alpha x 3 beta x 17 SNR x 100 channels = %d trajectories in total, sharded" % total, "strong", total
    if cfg == 5:
        B = batch or 8192
        return B, 128, 32, 76, "synthetic CDL mix Nt=128 Nr=32 Np=76, batch=%d per GPU" % B, "weak", B * world
    raise ValueError(cfg)


def make_batch(B, Nt, Nr, Np, rank=0, cfg=2):
    from score_based_channels_b200 import synth
    nuniq = min(B, 512)                                   # channel values do not change the work: tile a set of 512
    H = np.resize(synth.cdl_like_channels(nuniq, Nt, Nr, seed=4321 + 100000 * rank), (B, Nt, Nr))
    P = np.resize(synth.qpsk_pilots(nuniq, Nt, Np, seed=1234 + rank), (B, Np, Nt))
    snr = SNR_RANGE[np.arange(B) % len(SNR_RANGE)]
    nv = synth.snr_to_noise_var(snr, Nt).astype(np.float32)
    Y = synth.received_pilots(P, H, nv, seed=99 + rank)
    X0 = synth.cn01((B, Nt, Nr), np.random.default_rng(7 + rank))
    al = np.full(B, ALPHA_STEP, np.float32)
    be = np.full(B, BETA, np.float32)
    if cfg == 4:                                          # the (alpha, beta) cell of every trajectory
        cell = (np.arange(B) // (17 * 100)) % 12
        al = np.asarray([3e-11, 6e-11, 1e-10, 3e-10], np.float32)[cell // 3]
        be = np.asarray([0.1, 0.01, 0.001], np.float32)[cell % 3]
    return [np.ascontiguousarray(a) for a in (P, Y, X0, H, nv, al, be)]


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
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def cpu_arm(levels, B, max_seconds=25.0, Nt=64, Nr=16, Np=38):
    """The reference path restated on the CPU (oracle port), all host threads, on a bounded sample of the SAME workload:
    batch B (the GPU arm's batch), as many sigma levels as fit in `max_seconds`; the per-level cost does not depend on the
    level, so the full schedule is extrapolated."""
    from oracle import oracle as orc
    from score_based_channels_b200 import params
    sd = params.random_state(8, seed=1)
    cores = os.cpu_count() or 1
    orc.set_num_threads(cores)
    P, Y, X0, H, nv, al, be = make_batch(B, Nt, Nr, Np)
    net = orc.OracleNet(sd, 8, Nt, Nr)
    kw = dict(noise_var=nv, alpha_step=al, beta=be, sigma_end=SIGMA_END, steps_each=STEPS_EACH, seed=1)
    t0 = time.perf_counter()
    net.ald(P, Y, X0, H, level_begin=0, level_end=1, **kw)          # warm-up + calibration: 3 steps
    t_cal = (time.perf_counter() - t0) / STEPS_EACH
    n_lvl = int(max(1, min(levels, max_seconds / max(t_cal * STEPS_EACH, 1e-6))))
    t0 = time.perf_counter()
    net.ald(P, Y, X0, H, level_begin=0, level_end=n_lvl, **kw)
    dt = time.perf_counter() - t0
    t_step = dt / (n_lvl * STEPS_EACH)                              # seconds per Langevin step of the whole batch
    est_per_s = B / (t_step * NUM_LEVELS * STEPS_EACH)              # full-ALD estimates (2311 levels x 3 steps each)
    return {"value": est_per_s, "unit": "estimates/s", "cores": cores, "kind": "port", "batch": B,
            "core_seconds_per_estimate": cores * t_step * NUM_LEVELS * STEPS_EACH / B,
            "sample": "B=%d x %d levels x %d steps timed (%.1f s on %d threads), per-level cost extrapolated to %d levels"
                      % (B, n_lvl, STEPS_EACH, dt, cores, NUM_LEVELS), "ms_per_langevin_step": t_step * 1e3}


class Runner:
    """One (config, engine) workload resident on this rank's GPU."""

    def __init__(self, cfg, engine, levels, world, rank, dev, batch=None):
        import torch
        from score_based_channels_b200 import params
        from score_based_channels_b200.models import make_model
        self.torch, self.cfg, self.engine, self.levels, self.world, self.rank, self.dev = torch, cfg, engine, levels, world, rank, dev
        self.B, self.Nt, self.Nr, self.Np, self.desc, self.scaling, self.global_batch = workload(cfg, world, rank, batch)
        sd = params.random_state(8, seed=1)
        self.model = make_model(sd, ngf=8, precision=ENGINE_PRECISION[engine], Nt=self.Nt, Nr=self.Nr).to(dev)
        self.arrays = make_batch(self.B, self.Nt, self.Nr, self.Np, rank, cfg)
        self.host = [torch.from_numpy(a).pin_memory() for a in self.arrays]
        self.ids = torch.arange(rank * self.B, (rank + 1) * self.B, dtype=torch.int64, device=dev)
        self.nsteps_ald = levels * STEPS_EACH
        self.kw = dict(sigma_end=SIGMA_END, level_begin=0, level_end=levels, steps_each=STEPS_EACH, seed=2026)
        self.gather = None
        if world > 1:   # the single collective of the path: the per-rank NMSE logs (equal-sized shards only)
            self.gather = [torch.empty((self.nsteps_ald, self.B), dtype=torch.float32, device=dev) for _ in range(world)]
        self.pm = self.model.packed(self.Nt, self.Nr, dev)
        self.dres = [t.to(dev) for t in self.host]
        self.out_host = [torch.empty((self.B, self.Nt, self.Nr), dtype=torch.complex64).pin_memory(),
                         torch.empty((self.nsteps_ald, self.B), dtype=torch.float32).pin_memory()]

    def gpu_step(self, d):
        from score_based_channels_b200 import sampler
        import torch.distributed as dist
        P, Y, X0, H, nv, al, be = d
        X, nlog = sampler.ald_run(self.model, P, Y, X0, H, noise_var=nv, alpha_step=al, beta=be, sample_ids=self.ids, **self.kw)
        if self.gather is not None and self.scaling == "weak":
            dist.all_gather(self.gather, nlog)
        return X, nlog

    def resident(self):
        return self.gpu_step(self.dres)

    def e2e(self):
        d = [t.to(self.dev, non_blocking=True) for t in self.host]
        X, nlog = self.gpu_step(d)
        self.out_host[0].copy_(X, non_blocking=True)
        self.out_host[1].copy_(nlog, non_blocking=True)

    def e2e_host_abi(self):
        """The C-ABI host-buffer entry point sbc_ald_run_host on numpy arrays (what a non-torch caller uses)."""
        from score_based_channels_b200 import _lib
        P, Y, X0, H, nv, al, be = self.arrays
        if not hasattr(self, "_hx"):
            self._hx = X0.copy()
            self._hlog = np.empty((self.nsteps_ald, self.B), np.float32)
            self._hids = np.arange(self.rank * self.B, (self.rank + 1) * self.B, dtype=np.uint64)
        np.copyto(self._hx, X0)
        a = _lib.AldArgs(self.B, self.Nt, self.Nr, self.Np, 0, self.levels, STEPS_EACH, P.ctypes.data, Y.ctypes.data,
                         self._hx.ctypes.data, H.ctypes.data, nv.ctypes.data, al.ctypes.data, be.ctypes.data, SIGMA_END,
                         self._hlog.ctypes.data, 2026, self._hids.ctypes.data, None, None, None)
        _lib.check(_lib.lib().sbc_ald_run_host(self.pm.handle, C.byref(a)), "sbc_ald_run_host")

    def io_bytes(self):
        h2d = sum(t.numel() * t.element_size() for t in self.host)
        d2h = sum(t.numel() * t.element_size() for t in self.out_host)
        return int(h2d), int(d2h)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5])
    ap.add_argument("--engine", default="auto", choices=["auto", "1", "2"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--levels", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short runs of the other engine / the other configs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = args.config
    levels = args.levels or NUM_LEVELS
    B0, Nt, Nr, Np, desc, scaling, gbatch = workload(cfg, world, rank, args.batch)
    config = {"workload": "config %d: %s; %d sigma levels x %d steps, Nt=%d Nr=%d Np=%d" % (cfg, desc, levels, STEPS_EACH, Nt, Nr, Np),
              "global_batch": gbatch, "parallelism": "batch-sharded x%d" % world, "levels_run": levels,
              "l2_policy": "every timed step rewrites the kernel's whole working set (engine 1: shared memory; engine 2: a "
                           "per-CTA arena, 115 MB+ per launch > L2) and the e2e legs re-copy all inputs from the host"}

    if args.impl == "reference":
        if rank != 0:
            return
        times = []
        for i in range(args.warmup + args.steps):
            r = cpu_arm(levels, B0 if cfg != 4 else min(B0, 256), max_seconds=12.0, Nt=Nt, Nr=Nr, Np=Np)
            if i >= args.warmup:
                times.append(r)
        v = float(np.mean([r["value"] for r in times]))
        cb = dict(times[-1]); cb["value"] = v
        line = {"metric": METRIC, "value": v, "unit": "estimates/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * B0 / v, "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
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
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

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

    def est_per_s(total_traj, lv, n, ms):
        return total_traj * n * (lv / NUM_LEVELS) / (ms * 1e-3)      # full-ALD-equivalent estimates per second

    # engine choice (measured): the shared-memory-resident engine 1 wherever one sample fits an SM
    # (64x16, ngf 8); the tcgen05 engine 2 for the shapes that do not (128x32: engine 1 falls back to an L2 arena)
    if args.engine == "auto":
        engine = 1 if (Nt, Nr) == (64, 16) else 2
    else:
        engine = int(args.engine)

    R = Runner(cfg, engine, levels, world, rank, dev, args.batch)
    launches0 = R.pm.info().kernel_launches
    # ---- leg 1: inputs resident in HBM ("value") ----
    for _ in range(args.warmup):
        R.resident()
    with ClockSampler(local_rank) as cs:
        ms_res = timed(R.resident, args.steps)
    clocks = cs.summary()
    launches = R.pm.info().kernel_launches - launches0 - args.warmup
    # ---- leg 2: end to end through the public API with HOST buffers ("e2e") ----
    for _ in range(max(1, args.warmup // 3)):
        R.e2e()
    ms_e2e = timed(R.e2e, args.steps)
    # ---- leg 3: the same through the C-ABI host-buffer call (pageable numpy arrays) ----
    R.e2e_host_abi()
    ms_abi = timed(R.e2e_host_abi, max(1, args.steps // 2))
    n_abi = max(1, args.steps // 2)
    h2d, d2h = R.io_bytes()
    total = gbatch if scaling == "strong" else R.B * world
    value = est_per_s(total, levels, args.steps, ms_res)
    e2e = est_per_s(total, levels, args.steps, ms_e2e)
    e2e_abi = est_per_s(total, levels, n_abi, ms_abi)

    # ---- roofline of the dominant (only) kernel, timed inside the long step ----
    peaks, how = measured_peaks()
    info = R.pm.info()
    flops_per_launch = float(R.pm.conv_flops) * R.nsteps_ald * R.B       # dense conv FLOP, reference convention
    kern_s = ms_res * 1e-3 / args.steps                                  # one launch per step dominates the step
    achieved = flops_per_launch / kern_s / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    tr = TRAFFIC[engine]
    traffic = (tr["fixed"] + tr["per_level"] * levels) * R.B / tr["batch"]
    kernel_name = {1: "sbc_ald_kernel<smem arena, 3xTF32 mma.sync>", 2: "sbc2_ald_kernel (tcgen05.mma kind::f16 hi/lo split, TMEM accumulators)"}[engine]
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic,
                "traffic_note": "dram__bytes_read+write of the ncu capture in %s (fixed + per-level bytes at batch %d), scaled to this "
                                "launch's batch and schedule" % (tr["src"], tr["batch"]),
                "peak_source": how + ", sustained dense bf16", "kernel": kernel_name,
                "algorithmic_flops_per_launch": flops_per_launch, "engine": engine, "group_size": int(info.group_size),
                "ctas_per_sm": int(info.ctas_per_sm)}
    if engine == 1:
        # context, not the graded fraction: engine 1 contracts on legacy mma.sync TF32 (1024 FLOP/clk/SM) with three MMAs
        # per product, i.e. a ceiling of SMs x 1024 x clock / 3 algorithmic FLOP/s for this instruction path
        mhz = clocks.get("sm_mhz") or 1965.0
        legacy = int(info.num_sms) * 1024 * mhz * 1e6 / 3 / 1e12
        roofline["legacy_mma_tf32x3_ceiling_tflops"] = legacy
        roofline["frac_of_legacy_ceiling"] = achieved / legacy
        roofline["ncu"] = "profiles/r02b_ncu_engine1_2cta_raw.txt: tensor pipe 25.5 % active, issue slots 47.8 %, 2 CTAs/SM"

    extra = []
    if not args.no_extra and cfg == 2:
        # the other engine on the same workload, and configs 3 / 5 on a truncated schedule (see the module docstring)
        def short(c, e, lv, nrep=2, batch=None):
            try:
                r = Runner(c, e, lv, world, rank, dev, batch)
                r.resident()
                ms = timed(r.resident, nrep)
                tot = workload(c, world, rank, batch)[6] if workload(c, world, rank, batch)[5] == "strong" else r.B * world
                out = {"config": c, "engine": e, "precision": ENGINE_PRECISION[e], "batch_per_gpu": r.B, "levels_run": lv,
                       "value": est_per_s(tot, lv, nrep, ms), "unit": "estimates/s (full-ALD equivalent)",
                       "ms_per_step": ms / nrep, "workload": workload(c, world, rank, batch)[4]}
                del r
                torch.cuda.empty_cache()
                return out
            except Exception as ex:     # an extra must never take the headline line down
                return {"config": c, "engine": e, "error": str(ex)[:200]}
        extra.append(short(2, 2 if engine == 1 else 1, min(levels, 96)))
        extra.append(short(3, 2, min(levels, 12)))
        extra.append(short(3, 1, min(levels, 12)))
        extra.append(short(4, 1, min(levels, 6)))
        extra.append(short(4, 2, min(levels, 6)))
        extra.append(short(5, 2, min(levels, 3), batch=2368))
        extra.append(short(5, 1, min(levels, 3), batch=2368))

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_arm(levels, min(R.B, 256), Nt=Nt, Nr=Nr, Np=Np)
        line = {"metric": METRIC, "value": value, "unit": "estimates/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None,
                "dtype": {1: "tf32x3 (3xTF32 split, fp32-equivalent), f32 accumulate",
                          2: "fp16x2 (fp16 hi/lo split operands, 22 significant bits, fp32-equivalent), f32 accumulate"}[engine],
                "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e, "unit": "estimates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps, "api": "sampler.ald_run on pinned host tensors"},
                "e2e_host_abi": {"value": e2e_abi, "unit": "estimates/s", "ms_per_step": ms_abi / n_abi,
                                 "api": "sbc_ald_run_host (C ABI, pageable numpy buffers, preallocated device workspace)"},
                "roofline": roofline, "cpu_baseline": cpu, "extra": extra}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
