#!/usr/bin/env python3
"""BASELINE config 5 shape (Nt=128, Nr=32, Np=76, synthetic channels): functional + timing probe of the global-arena
path (the arena of this shape, ~0.9 MB per sample, does not fit shared memory).  Prints estimates/s extrapolated to the
full schedule from a few levels."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from score_based_channels_b200 import params, sampler, synth  # noqa: E402
from score_based_channels_b200.models import make_model  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0")
sd = params.random_state(8, seed=1)
Nt, Nr, Np = 128, 32, 76
m = make_model(sd, ngf=8, Nt=Nt, Nr=Nr).to(dev)
H = synth.cdl_like_channels(B, Nt, Nr)
P = synth.qpsk_pilots(B, Nt, Np)
nv = float(synth.snr_to_noise_var(10.0, Nt))
Y = synth.received_pilots(P, H, nv)
X0 = synth.cn01((B, Nt, Nr), np.random.default_rng(3))
t = [torch.from_numpy(a).to(dev) for a in (P, Y, X0, H)]
kw = dict(noise_var=nv, alpha_step=3e-11, beta=0.01, sigma_end=2.599515446446343e-4, level_begin=0, level_end=levels,
          steps_each=3, seed=1)
sampler.ald_run(m, *t, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
X, nlog = sampler.ald_run(m, *t, **kw)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
info = m.packed(Nt, Nr, dev).info()
print("config-5 shape Nt=%d Nr=%d B=%d: arena_in_smem=%d arena=%.0f KB; %d levels x 3 steps in %.1f ms -> %.3f full-ALD "
      "estimates/s (extrapolated to 2311 levels); NMSE finite: %s"
      % (Nt, Nr, B, info.arena_in_smem, info.arena_bytes / 1024, levels, ms, B / (ms * 1e-3 * 2311 / levels),
         bool(torch.isfinite(nlog).all())))
