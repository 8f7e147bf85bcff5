#!/usr/bin/env python3
"""Generates the golden vectors under tests/golden/ from the REFERENCE's own Python modules.

Runs only in the build container (needs /root/reference, which does not exist on the GPU box);
the produced .npz files are committed and are what the oracle (oracle/sbc_oracle.c) and the CUDA
path are pinned against.  The network is the unmodified ``ncsnv2.models.ncsnv2.NCSNv2Deepest``;
the sampler loop below is ``src/score_based_channels/test_score.py:118-171`` with ``.cuda()``
dropped and the ``torch.randn_like`` draws recorded so that other implementations can replay them.

Weights are NOT stored: every case uses ``params.random_state(ngf, seed=...)`` which is
reproducible from the seed on any machine (numpy default_rng).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")

from score_based_channels_b200 import dotmap_shim  # noqa: E402

dotmap_shim.install()
from dotmap import DotMap  # noqa: E402

from ncsnv2.models.ncsnv2 import NCSNv2, NCSNv2Deeper, NCSNv2Deepest  # noqa: E402  (the reference modules)
from score_based_channels_b200 import params, synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SIGMA_BEGIN, SIGMA_END, L = 27.77, 2.599515446446343e-4, 2311


REF_CLASS = {"deepest": NCSNv2Deepest, "deeper": NCSNv2Deeper, "ncsnv2": NCSNv2}


def ref_model(ngf, wseed, H=64, W=16, arch="deepest"):
    cfg = DotMap()
    cfg.device = "cpu"
    cfg.model.ngf = ngf
    cfg.model.num_classes = L
    cfg.model.normalization = "InstanceNorm++"
    cfg.model.nonlinearity = "elu"
    cfg.model.sigma_dist = "geometric"
    cfg.model.sigma_begin = SIGMA_BEGIN
    cfg.model.sigma_end = SIGMA_END
    cfg.data.channels = 2
    cfg.data.image_size = [W, H]
    m = REF_CLASS[arch](cfg)
    sd = params.random_state(ngf, seed=wseed, sigma_begin=SIGMA_BEGIN, sigma_end=SIGMA_END, num_classes=L, arch=arch)
    print(m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}))
    return m.eval(), cfg


def golden_forward(name, ngf, wseed, H, W, labels, scales, strided=False, arch="deepest"):
    m, _ = ref_model(ngf, wseed, H, W, arch)
    torch.manual_seed(100 + wseed)
    B = len(labels)
    x = torch.randn(B, 2, H, W) * torch.tensor(scales).view(B, 1, 1, 1)
    y = torch.tensor(labels)
    with torch.no_grad():
        out = m(x, y)
    np.savez_compressed(os.path.join(OUT, name), ngf=ngf, wseed=wseed, x=x.numpy(), y=y.numpy(), out=out.numpy(), arch=arch)
    print(name, "out absmax per sample", out.abs().amax(dim=(1, 2, 3)))


def reference_ald(m, cfg, val_P, val_H, val_Y, init_val_H, local_noise, alpha_step, beta_noise, levels,
                  steps_each):
    """test_score.py:126-171, for one SNR point; records the noise draws and the state after every step."""
    current = init_val_H.clone()
    y = val_Y
    forward = val_P
    forward_h = torch.conj(torch.transpose(val_P, -1, -2))
    oracle = val_H
    xs, noises, nmse = [], [], []
    for step_idx in levels:
        current_sigma = m.sigmas[step_idx].item()
        labels = (torch.ones(init_val_H.shape[0]) * step_idx).long()
        alpha = alpha_step * (current_sigma / cfg.model.sigma_end) ** 2
        for inner_idx in range(steps_each):
            current_real = torch.view_as_real(current).permute(0, 3, 1, 2)
            with torch.no_grad():
                score = m(current_real, labels)
            score = torch.view_as_complex(score.permute(0, 2, 3, 1).contiguous())
            meas_grad = torch.matmul(forward_h, torch.matmul(forward, current) - y)
            eps = torch.randn_like(current)
            noises.append(eps.numpy().copy())
            grad_noise = np.sqrt(2 * alpha * beta_noise) * eps
            current = current + alpha * (score - meas_grad / (local_noise / 2. + current_sigma ** 2)) + grad_noise
            nmse.append((torch.sum(torch.square(torch.abs(current - oracle)), dim=(-1, -2)) /
                         torch.sum(torch.square(torch.abs(oracle)), dim=(-1, -2))).numpy())
            xs.append(current.numpy().copy())
    return np.stack(xs), np.stack(noises), np.stack(nmse)


def golden_ald(name, ngf, wseed, B, Np, snr_db, levels, steps_each, alpha_step, beta_noise, Nt=64, Nr=16,
               init_sigma=None):
    m, cfg = ref_model(ngf, wseed, Nt, Nr)
    Hc = synth.cdl_like_channels(B, Nt, Nr, seed=4321)
    P = synth.qpsk_pilots(B, Nt, Np, seed=1234)
    local_noise = float(synth.snr_to_noise_var(snr_db, Nt))
    torch.manual_seed(7 + wseed)
    val_P, val_H = torch.from_numpy(P), torch.from_numpy(Hc)
    init_val_H = torch.randn_like(val_H)
    if init_sigma is not None:  # mid-trajectory start: truth + sigma * CN(0,1)
        init_val_H = val_H + init_sigma * init_val_H
    val_Y = torch.matmul(val_P, val_H)
    val_Y = val_Y + np.sqrt(local_noise) * torch.randn_like(val_Y)
    xs, noises, nmse = reference_ald(m, cfg, val_P, val_H, val_Y, init_val_H, local_noise, alpha_step, beta_noise,
                                     levels, steps_each)
    np.savez_compressed(os.path.join(OUT, name), ngf=ngf, wseed=wseed, P=P, H=Hc, Y=val_Y.numpy(),
                        X0=init_val_H.numpy(), noise_var=local_noise, alpha_step=alpha_step, beta=beta_noise,
                        levels=np.asarray(levels), steps_each=steps_each, sigma_end=SIGMA_END, xs=xs,
                        ext_noise=noises, nmse=nmse)
    print(name, "nmse per step (mean over batch)", nmse.mean(1))


def golden_real_checkpoint(name="real_ckpt_forward.npz"):
    """The SHIPPED checkpoint (pretrained_models/score-deepest-cdl-c.pt) on the first shipped CDL-C validation
    channels at six noise levels: reference module output in fp32 and in fp64.  With trained weights the
    output at small sigma is an ill-conditioned difference (fp32 vs fp64 differ by up to 3.5e-4 relative), so
    parity tests bound the error of an implementation by a multiple of the reference's own fp32 error."""
    from score_based_channels_b200 import entry_common as ec, hdf5_min
    contents = ec.load_checkpoint("/root/reference/pretrained_models/score-deepest-cdl-c.pt")
    h = hdf5_min.loadmat_v73("/root/reference/sample_data/CDL-C_Nt64_Nr16_ULA0.50_seed4321.mat")["output_h"]
    h = h[:6, 0].astype(np.complex64)
    Hn = np.conj(np.transpose(h, (0, 2, 1))) / 0.363263
    labels = np.array([0, 500, 1000, 1500, 2000, 2310])
    sig = contents["model_state"]["sigmas"].numpy()
    rng = np.random.default_rng(0)
    x = Hn + sig[labels][:, None, None] * ((rng.standard_normal(Hn.shape) + 1j * rng.standard_normal(Hn.shape)) / np.sqrt(2))
    xr = np.stack([x.real, x.imag], 1).astype(np.float32)
    cfg = contents["config"]
    cfg.device = "cpu"
    m = NCSNv2Deepest(cfg)
    print(m.load_state_dict(contents["model_state"]))
    m.eval()
    with torch.no_grad():
        o32 = m(torch.from_numpy(xr), torch.from_numpy(labels)).numpy()
        torch.set_default_dtype(torch.float64)      # MSFBlock allocates its accumulator with the default dtype
        o64 = m.double()(torch.from_numpy(xr).double(), torch.from_numpy(labels)).numpy()
        torch.set_default_dtype(torch.float32)
    np.savez_compressed(os.path.join(OUT, name), x=xr, y=labels, out32=o32, out64=o64)
    rel = [float(np.linalg.norm(o32[b] - o64[b]) / np.linalg.norm(o64[b])) for b in range(6)]
    print(name, "reference fp32 vs fp64 relative error per sample", rel)


def golden_dsm(name="dsm_val.npz", ngf=8, wseed=1, B=4, H=64, W=16, anneal_power=2.0, seed=77):
    """anneal_dsm_score_estimation (ncsnv2/losses/dsm.py:6-32) evaluated under no_grad as train_score.py:170-185 does for the
    validation loss: labels given, the randn_like draw recovered by re-seeding."""
    from ncsnv2.losses.dsm import anneal_dsm_score_estimation
    m, _ = ref_model(ngf, wseed, H, W)
    rng = np.random.default_rng(seed)
    Hc = synth.cdl_like_channels(B, H, W, seed=seed)
    samples = torch.from_numpy(np.stack([Hc.real, Hc.imag], 1).astype(np.float32))      # 'H_herm' layout [B,2,Nt,Nr]
    labels = torch.from_numpy(rng.integers(0, L, B))
    labels[0], labels[1] = 0, L - 1
    with torch.no_grad():
        torch.manual_seed(seed)
        loss = anneal_dsm_score_estimation(m, samples, m.sigmas, labels, anneal_power)
        torch.manual_seed(seed)
        z = torch.randn_like(samples)
        # per-sample terms, restated from the reference lines to store more than one number
        us = m.sigmas[labels].view(B, 1, 1, 1)
        noise = z * us
        scores = m(samples + noise, labels)
        per = 0.5 * ((scores.view(B, -1) - (-1 / us ** 2 * noise).view(B, -1)) ** 2).sum(-1) * us.squeeze() ** anneal_power
        assert torch.allclose(per.mean(), loss, rtol=1e-6), (per.mean(), loss)
    np.savez_compressed(os.path.join(OUT, name), ngf=ngf, wseed=wseed, samples=samples.numpy(), labels=labels.numpy(),
                        z=z.numpy(), anneal_power=anneal_power, per_sample=per.numpy(), loss=float(loss))
    print(name, "loss", float(loss), "per sample", per.numpy())


if __name__ == "__main__":
    golden_dsm()
    if "--dsm-only" in sys.argv:
        sys.exit(0)
    golden_real_checkpoint()
    golden_forward("forward_ngf8.npz", 8, 1, 64, 16, [0, 1000, 2310, 17, 2000], [27.0, 0.5, 1.0, 10.0, 0.05])
    golden_forward("forward_ngf8_32x8.npz", 8, 2, 32, 8, [5, 1500], [20.0, 1.0])
    golden_forward("forward_ngf16.npz", 16, 3, 64, 16, [0, 2310], [27.0, 1.0])
    # the other two score nets of the reference (ncsnv2.py:11-195)
    golden_forward("forward_deeper_ngf8.npz", 8, 5, 64, 16, [0, 1200, 2310], [27.0, 1.0, 0.05], arch="deeper")
    golden_forward("forward_ncsnv2_ngf8.npz", 8, 6, 64, 16, [0, 1200, 2310], [27.0, 1.0, 0.05], arch="ncsnv2")
    # BASELINE config 1: batch 4, 2 sigma levels x 3 steps, SNR 0 dB, alpha0 = 3e-11, beta = 0.01
    golden_ald("ald_cfg1.npz", 8, 1, B=4, Np=38, snr_db=0.0, levels=[0, 1], steps_each=3, alpha_step=3e-11,
               beta_noise=0.01)
    # mid-trajectory levels (sigma ~ 0.18, where the score and the data term are both active), SNR 20 dB
    golden_ald("ald_mid.npz", 8, 1, B=2, Np=38, snr_db=20.0, levels=[1000, 1001], steps_each=3, alpha_step=3e-10,
               beta_noise=0.01, init_sigma=0.18)
