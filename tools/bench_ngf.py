#!/usr/bin/env python3
"""Short ALD timing at other network widths (train_score.py:43 default is ngf=32): python tools/bench_ngf.py [ngf] [B] [levels]"""
import os, sys, time
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from score_based_channels_b200 import params, sampler, synth
from score_based_channels_b200.models import make_model
ngf = int(sys.argv[1]) if len(sys.argv) > 1 else 32
B = int(sys.argv[2]) if len(sys.argv) > 2 else 592
levels = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
sd = params.random_state(ngf, seed=1)
Nt, Nr, Np = 64, 16, 38
H = synth.cdl_like_channels(B, Nt, Nr); P = synth.qpsk_pilots(B, Nt, Np)
nv = float(synth.snr_to_noise_var(10.0, Nt)); Y = synth.received_pilots(P, H, nv)
X0 = synth.cn01((B, Nt, Nr), np.random.default_rng(3))
d = [torch.from_numpy(a).to(dev) for a in (P, Y, X0, H)]
kw = dict(noise_var=nv, alpha_step=3e-11, beta=0.01, sigma_end=2.599515446446343e-4, level_begin=0, level_end=levels, steps_each=3, seed=11)
for prec in ("tf32x3", "fp16x2"):
    m = make_model(sd, ngf=ngf, precision=prec).to(dev)
    sampler.ald_run(m, *d, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sampler.ald_run(m, *d, **kw); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    info = m.packed(Nt, Nr, dev).info()
    print("ngf=%d %s (engine %d, %d CTAs/SM, arena in smem: %d): B=%d, %d levels x 3: %.1f ms -> %.2f full-ALD estimates/s"
          % (ngf, prec, info.engine, info.ctas_per_sm, info.arena_in_smem, B, levels, ms, B * (levels / 2311.0) / (ms * 1e-3)))
