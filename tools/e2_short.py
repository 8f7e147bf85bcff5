#!/usr/bin/env python3
"""One short engine-2 ALD launch (for ncu captures): python tools/e2_short.py [batch] [levels]"""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import __graft_entry__ as ge
ge.build()
from score_based_channels_b200 import params, sampler, synth
from score_based_channels_b200.models import make_model
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 2
prec = sys.argv[3] if len(sys.argv) > 3 else "fp16x2"
dev = torch.device("cuda:0")
sd = params.random_state(8, seed=1)
m = make_model(sd, ngf=8, precision=prec).to(dev)
Nt, Nr, Np = 64, 16, 38
H = synth.cdl_like_channels(B, Nt, Nr); P = synth.qpsk_pilots(B, Nt, Np)
nv = float(synth.snr_to_noise_var(10.0, Nt)); Y = synth.received_pilots(P, H, nv)
X0 = synth.cn01((B, Nt, Nr), np.random.default_rng(3))
d = [torch.from_numpy(a).to(dev) for a in (P, Y, X0, H)]
kw = dict(noise_var=nv, alpha_step=3e-11, beta=0.01, sigma_end=2.599515446446343e-4, level_begin=0, level_end=levels, steps_each=3, seed=11)
for _ in range(2):
    sampler.ald_run(m, *d, **kw)
torch.cuda.synchronize()
print("done")
