#!/usr/bin/env python3
"""Per-op cycle profile of the engine-2 kernel (clock64 stamps of CTA 0, second Langevin step).
   python tools/e2_profile.py [--batch 256] [--out profiles/r02_e2_ops_profile.txt]"""
import argparse
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import __graft_entry__ as ge  # noqa: E402

ge.build()
from score_based_channels_b200 import _lib, params, sampler, synth  # noqa: E402
from score_based_channels_b200.models import make_model  # noqa: E402

KINDS = {0: "affine", 1: "conv", 2: "norm_elu", 3: "elu", 4: "maxpool5", 5: "upacc", 6: "pool2", 7: "epilogue"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    sd = params.random_state(8, seed=1)
    m = make_model(sd, ngf=8, precision="fp16x2").to(dev)
    B, Nt, Nr, Np = args.batch, 64, 16, 38
    H = synth.cdl_like_channels(B, Nt, Nr)
    P = synth.qpsk_pilots(B, Nt, Np)
    nv = float(synth.snr_to_noise_var(10.0, Nt))
    Y = synth.received_pilots(P, H, nv)
    X0 = synth.cn01((B, Nt, Nr), np.random.default_rng(3))
    d = [torch.from_numpy(a).to(dev) for a in (P, Y, X0, H)]
    kw = dict(noise_var=nv, alpha_step=3e-11, beta=0.01, sigma_end=2.599515446446343e-4, level_begin=0, level_end=2,
              steps_each=3, seed=11)
    pm = m.packed(Nt, Nr, dev)
    sampler.ald_run(m, *d, **kw)
    n_ops = pm.info().n_ops
    stamps = torch.zeros(n_ops + 2 + 3 * 64 * 4, dtype=torch.int64, device=dev)
    _lib.check(_lib.lib().sbc_set_profile_buffer(pm.handle, stamps.data_ptr()), "prof")
    sampler.ald_run(m, *d, **kw)
    torch.cuda.synchronize()
    _lib.check(_lib.lib().sbc_set_profile_buffer(pm.handle, None), "prof")
    st = stamps.cpu().numpy()
    info = pm.info()
    names = pm.op_names()
    kinds = [_lib.lib().sbc_op_kind(pm.handle, i) for i in range(n_ops)]
    dur = np.diff(st[:n_ops + 1])
    lines = ["engine-2 per-op profile: B=%d, group size S=%d, %d CTAs/SM; cycles of CTA 0, second step"
             % (B, info.group_size, info.ctas_per_sm),
             "network: %d clk, Langevin tail: %d clk, total step: %d clk" % (st[n_ops] - st[0], st[n_ops + 1] - st[n_ops], st[n_ops + 1] - st[0])]
    bykind = {}
    for i in range(n_ops):
        bykind.setdefault(KINDS[kinds[i]], []).append(dur[i])
    for k, v in sorted(bykind.items(), key=lambda kv: -sum(kv[1])):
        lines.append("  %-9s n=%3d total %8d clk (%.1f%%)  mean %6.0f  min %6d  max %6d"
                     % (k, len(v), sum(v), 100.0 * sum(v) / (st[n_ops + 1] - st[0]), np.mean(v), min(v), max(v)))
    lines.append("per op:")
    for i in range(n_ops):
        lines.append("  %3d %-9s %7d  %s" % (i, KINDS[kinds[i]], dur[i], names[i]))
    top = int(os.environ.get("SBC2_TRACE_OP", "-1"))
    if top >= 0:
        tr = st[n_ops + 2:].reshape(3, 64, 4)
        t0 = st[top]
        if kinds[top] == 2:
            lines.append("norm trace of op %d (%s): start, loads+sum, shuffles, sync1, cmean+sync2, pass2+sync3, coef+sync4, apply; cycles from op start: %s; op end %d"
                         % (top, names[top], [int(v - t0) for v in st[n_ops + 2:n_ops + 10]], st[top + 1] - t0))
        lines.append("intra-conv trace of op %d (%s), cycles relative to the op start:" % (top, names[top]))
        lines.append("  tile | producer: loop-top  stage-free  issued | mma: loop-top  stage-full  acc-free  issued | epilogue(warp0): loop-top  acc-full  stored")
        for t in range(64):
            if tr[0, t, 0] == 0 and tr[1, t, 0] == 0:
                break
            r = lambda a: "%7d" % (a - t0) if a else "      -"
            lines.append("  %4d | %s %s %s | %s %s %s %s | %s %s %s" % (t, r(tr[0, t, 0]), r(tr[0, t, 1]), r(tr[0, t, 2]), r(tr[1, t, 0]), r(tr[1, t, 1]),
                                                                 r(tr[1, t, 2]), r(tr[1, t, 3]), r(tr[2, t, 0]), r(tr[2, t, 1]), r(tr[2, t, 2])))
        lines.append("  op end (next op start): %d" % (st[top + 1] - t0))
    txt = "\n".join(lines)
    print(txt)
    if args.out:
        with open(args.out, "w") as f:
            f.write(txt + "\n")


if __name__ == "__main__":
    main()
