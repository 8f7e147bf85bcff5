#!/usr/bin/env python3
"""Engine-2 (tcgen05) bring-up check on a GPU box: (1) one forward against the reference golden, (2) if it is off, a
tensor-by-tensor comparison of the kernel's group arena with the CPU plan emulation (tests/emu/emu2.cpp) to name the
first op that diverges, (3) a short timed ALD run.   python tools/e2_check.py [--levels 8] [--batch 256]"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import __graft_entry__ as ge  # noqa: E402

ge.build()
from score_based_channels_b200 import _lib, params, sampler, synth  # noqa: E402
from score_based_channels_b200.models import make_model  # noqa: E402


def decode(arena, t, geo, S):
    G = geo[t.level]
    h, w, wp, pps, lead, npx = int(G[0]), int(G[1]), int(G[4]), int(G[6]), int(G[7]), int(G[8])
    raw = arena[t.off:t.off + t.bytes]
    ss, yy, xx = np.meshgrid(np.arange(S), np.arange(h), np.arange(w), indexing="ij")
    q = lead + ss * pps + yy * wp + xx
    res = np.zeros((S, t.C, h, w), np.float32)
    if t.fmt == 0:
        a = raw.view(np.float32).reshape(t.C // 4, npx, 4)
        for c in range(t.C):
            res[:, c] = a[c // 4, q, c % 4]
    else:
        a = raw.view(np.float16).reshape(t.C // 8, 2, npx, 8).astype(np.float32)
        for c in range(t.C):
            res[:, c] = a[c // 8, 0, q, c % 8] + a[c // 8, 1, q, c % 8]
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--levels", type=int, default=8)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--skip-timing", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    g = np.load(os.path.join(REPO, "tests", "golden", "forward_ngf8.npz"))
    sd = params.random_state(8, seed=int(g["wseed"]))
    model = make_model(sd, ngf=8, precision="fp16x2").to(dev)
    x = torch.from_numpy(g["x"]).to(dev)
    y = torch.from_numpy(g["y"]).to(dev)
    out = model(x, y).cpu().numpy()
    torch.cuda.synchronize()
    rel = [float(np.linalg.norm(out[b] - g["out"][b]) / np.linalg.norm(g["out"][b])) for b in range(out.shape[0])]
    print("[e2] forward rel err vs reference golden:", ["%.2e" % r for r in rel], flush=True)
    pm = model.packed(64, 16, dev)
    info = pm.info()
    print("[e2] engine %d ctas/sm %d smem/cta %d group %d n_ops %d" % (info.engine, info.ctas_per_sm,
                                                                      info.smem_bytes_per_cta, info.group_size, info.n_ops))
    if max(rel) > 2e-5 or not np.isfinite(max(rel)):
        import test_emulation2 as T
        S = 2
        emu = T.Emu2(T.load_emu2(), sd, 8, 64, 16)
        ab, tens, geo = pm.debug_plan(S, reuse=False)
        ab2, tens2, geo2 = emu.plan(S, reuse=False)
        assert ab == ab2 and (geo == geo2).all()
        xs = np.ascontiguousarray(g["x"][:S])
        ea = np.zeros(ab, np.uint8)
        eo = np.empty_like(xs)
        emu.lib.emu2_forward(emu.h, S, 0, xs.ctypes.data, eo.ctypes.data, ea.ctypes.data, -1)
        ga = torch.zeros(ab, dtype=torch.uint8, device=dev)
        _lib.check(_lib.lib().sbc_debug_run(pm.handle, torch.from_numpy(xs).to(dev).data_ptr(), S, 0, ga.data_ptr(), None),
                   "sbc_debug_run")
        ga = ga.cpu().numpy()
        names = pm.op_names()
        bad = 0
        for t in sorted(tens, key=lambda t: t.born):
            if t.fmt == 2:
                continue
            a, b = decode(ga, t, geo, S), decode(ea, t, geo, S)
            err = float(np.abs(a - b).max()) / (float(np.abs(b).max()) + 1e-30)
            flag = "" if err < 1e-4 else "   <-- MISMATCH"
            if flag or bad < 3:
                print("[e2] tensor %-8s fmt %d lvl %d C %3d born at op %3d (%s): rel err %.3e%s"
                      % (t.name.decode(), t.fmt, t.level, t.C, t.born, names[min(t.born, len(names) - 1)], err, flag))
            if flag:
                bad += 1
                if bad >= 6:
                    break
        return 1
    if args.skip_timing:
        return 0
    B, Nt, Nr, Np = args.batch, 64, 16, 38
    H = synth.cdl_like_channels(B, Nt, Nr)
    P = synth.qpsk_pilots(B, Nt, Np)
    nv = float(synth.snr_to_noise_var(10.0, Nt))
    Y = synth.received_pilots(P, H, nv)
    X0 = synth.cn01((B, Nt, Nr), np.random.default_rng(3))
    d = [torch.from_numpy(a).to(dev) for a in (P, Y, X0, H)]
    kw = dict(noise_var=nv, alpha_step=3e-11, beta=0.01, sigma_end=2.599515446446343e-4, level_begin=0,
              level_end=args.levels, steps_each=3, seed=11)
    for prec in ("fp16x2", "tf32x3"):
        m = make_model(sd, ngf=8, precision=prec).to(dev)
        sampler.ald_run(m, *d, **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        X, nlog = sampler.ald_run(m, *d, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("[e2] %s: B=%d, %d levels x 3: %.1f ms -> %.1f full-ALD estimates/s (%.1f us per network evaluation of the batch)"
              % (prec, B, args.levels, dt * 1e3, B * (args.levels / 2311.0) / dt, dt * 1e6 / (args.levels * 3)), flush=True)
        if prec == "fp16x2":
            X2, n2 = X.cpu().numpy(), nlog.cpu().numpy()
        else:
            print("[e2] fp16x2 vs tf32x3: max |dX| / max|X| = %.3e, NMSE-log rel diff = %.3e"
                  % (np.abs(X2 - X.cpu().numpy()).max() / np.abs(X2).max(), np.abs(n2 - nlog.cpu().numpy()).max() / np.abs(n2).max()))
    return 0


if __name__ == "__main__":
    sys.exit(main())
