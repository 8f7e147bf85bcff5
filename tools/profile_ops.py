#!/usr/bin/env python3
"""Per-op cycle profile of the fused kernel (clock64 stamps of CTA 0) + a plain timed forward.
Usage on the GPU box:  python tools/profile_ops.py [B] > gpurun_out/ops_profile.txt"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from score_based_channels_b200 import _lib, params, program, sampler, synth  # noqa: E402
from score_based_channels_b200.models import make_model  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
PREC = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
dev = torch.device("cuda:0")
sd = params.random_state(8, seed=1)
m = make_model(sd, ngf=8, precision=PREC).to(dev)
pm = m.packed(64, 16, dev)
prog = pm.prog
x = torch.randn(B, 2, 64, 16, device=dev)
y = torch.zeros(B, dtype=torch.long, device=dev)
for _ in range(3):
    m(x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    m(x, y)
e1.record()
torch.cuda.synchronize()
print("precision %s forward B=%d: %.3f ms per launch" % (PREC, B, e0.elapsed_time(e1) / 10))

stamps = torch.zeros(6 * len(prog.ops) + 2 + 2 * max(B, 1024), dtype=torch.int64, device=dev)
_lib.check(_lib.lib().sbc_set_profile_buffer(pm.handle, stamps.data_ptr()), "prof")
P = torch.from_numpy(synth.qpsk_pilots(B, 64, 38)).to(dev)
H = torch.from_numpy(synth.cdl_like_channels(B)).to(dev)
Y = torch.matmul(P, H)
X0 = torch.randn_like(H)
sampler.ald_run(m, P, Y, X0, H, noise_var=1.0, alpha_step=3e-11, beta=0.01, level_begin=0, level_end=1, steps_each=2)
torch.cuda.synchronize()
_lib.check(_lib.lib().sbc_set_profile_buffer(pm.handle, None), "prof")
full = stamps.cpu().numpy()
st = full[:len(prog.ops) + 2]
sub = full[len(prog.ops) + 2:5 * len(prog.ops) + 2].reshape(len(prog.ops), 4)
tw = full[5 * len(prog.ops) + 2:6 * len(prog.ops) + 2]
cta = full[6 * len(prog.ops) + 2:6 * len(prog.ops) + 2 + 2 * min(B, 2 * 148)].reshape(-1, 2)
d = np.diff(st)
tot = st[-1] - st[0]
print("CTA0 second step (warm): %d cycles total; network %d, langevin %d" % (tot, st[-2] - st[0], st[-1] - st[-2]))
kinds = {0: "affine", 1: "conv", 2: "norm_elu", 3: "elu", 4: "maxpool5", 5: "upacc", 6: "conv_mma", 7: "spill", 8: "fill"}
by_kind = {}
rows = []
for i, op in enumerate(prog.ops):
    c = int(d[i])
    key = kinds[op.kind]
    if op.kind == 6:
        key = "conv_mma %dx%d c%d->%d k%d d%d%s ks%d" % (op.h, op.w, op.cin, op.cout, op.ksize, op.dil,
                                                        " pool" if op.flags & 1 else "", op.ks)
    else:
        key = "%s %dx%d c%d" % (key, op.h, op.w, op.cin)
    by_kind.setdefault(key, [0, 0])
    by_kind[key][0] += c
    by_kind[key][1] += 1
    ss = sub[i]
    phases = ""
    if ss[0] > 0 and ss[3] > 0:   # start->prologue | ->loop entry | loop | epilogue | barrier+tail
        e = int(st[i + 1])
        phases = " | pro %d ent %d loop %d epi %d bar %d" % (ss[0] - st[i], max(ss[1] - ss[0], 0), max(ss[2] - ss[1], 0),
                                                          ss[3] - max(ss[2], ss[0]), e - ss[3])
    if tw[i] > 0:
        phases += " | fetch+wait %d" % (tw[i] - st[i])
    rows.append((i, op.name, key + phases, c))
cyc, sm = cta[:, 0], cta[:, 1]
occ = np.bincount(sm.astype(np.int64), minlength=148)
two = cyc[occ[sm] == 2]; one = cyc[occ[sm] == 1]
print("per-CTA total cycles (whole launch: 2 steps): all CTAs min %d median %d max %d; on SMs with 2 CTAs: n=%d median %d max %d; alone on an SM: n=%d median %d"
      % (cyc.min(), np.median(cyc), cyc.max(), two.size, np.median(two) if two.size else 0, two.max() if two.size else 0, one.size, np.median(one) if one.size else 0))
print("\n== by op class (cycles, count, cycles/op, share of network)")
net = st[-2] - st[0]
for k, (c, n) in sorted(by_kind.items(), key=lambda kv: -kv[1][0]):
    print("%-48s %9d %4d %8.0f %6.2f%%" % (k, c, n, c / n, 100.0 * c / net))
print("\n== every op")
for r in rows:
    print("%3d %-36s %8d  %s" % (r[0], r[1], r[3], r[2]))
