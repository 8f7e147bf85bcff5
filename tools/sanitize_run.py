#!/usr/bin/env python3
"""Tiny ALD + forward run for compute-sanitizer (memcheck / racecheck / synccheck): B=3 samples, 1 level x 2 steps,
both precision modes, checked against the oracle so that a sanitizer-clean run is also a correct one."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc  # noqa: E402
from score_based_channels_b200 import params, sampler, synth  # noqa: E402
from score_based_channels_b200.models import make_model  # noqa: E402

dev = torch.device("cuda:0")
sd = params.random_state(8, seed=1)
B, Nt, Nr, Np = 3, 64, 16, 38
H = synth.cdl_like_channels(B, Nt, Nr)
P = synth.qpsk_pilots(B, Nt, Np)
nv = float(synth.snr_to_noise_var(0.0, Nt))
Y = synth.received_pilots(P, H, nv)
X0 = synth.cn01((B, Nt, Nr), np.random.default_rng(3))
kw = dict(noise_var=nv, alpha_step=3e-11, beta=0.01, sigma_end=2.599515446446343e-4, level_begin=0, level_end=1,
          steps_each=2, seed=11)
Xo, nlo = orc.OracleNet(sd, 8, Nt, Nr).ald(P, Y, X0, H, **kw)
for prec in (sys.argv[1:] or ["tf32x3", "tf32"]):
    model = make_model(sd, ngf=8, precision=prec).to(dev)
    X, nlog = sampler.ald_run(model, *(torch.from_numpy(a).to(dev) for a in (P, Y, X0, H)), **kw)
    torch.cuda.synchronize()
    err = np.abs(X.cpu().numpy() - Xo).max() / np.abs(Xo).max()
    print("[sanitize_run] %s: max |X - oracle| / max|X| = %.3e" % (prec, err))
    assert err < (5e-3 if prec == "tf32" else 1e-4)
    x = torch.randn(2, 2, Nt, Nr, device=dev)
    out = model(x, torch.tensor([0, 2310], device=dev))
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
print("[sanitize_run] done")
