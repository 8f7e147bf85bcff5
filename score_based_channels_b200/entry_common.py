"""Shared host logic of the two reference-compatible entry points (test_score / tune_hparams_score)."""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np
import torch

from . import dist as sdist
from . import dotmap_shim, sampler
from .ncsnv2 import NCSNv2Deepest


def load_checkpoint(target_file: str, model_name: Optional[str] = None):
    """``contents = torch.load(target_file)`` of reference test_score.py:33-36.  The file is a pickle that
    embeds a ``dotmap.DotMap`` config; the bundled shim stands in for the (absent) dotmap package.

    When ``target_file`` is missing the reference raises.  Two substitutions are allowed here, both announced:
    an explicit ``$SBC_CKPT`` override, and -- only when the requested model IS the shipped one (``model_name``
    'CDL-C') -- the shipped ``score-deepest-cdl-c.pt`` fixture.  Anything else raises FileNotFoundError, so results
    are never written under another model's name."""
    dotmap_shim.install()
    if not os.path.exists(target_file):
        alt = [os.environ.get("SBC_CKPT", "")]
        if model_name == "CDL-C":
            alt += ["./fixtures_local/score-deepest-cdl-c.pt", "./pretrained_models/score-deepest-cdl-c.pt"]
        for a in alt:
            if a and os.path.exists(a):
                print("checkpoint %s not found; using %s" % (target_file, a))
                target_file = a
                break
        else:
            raise FileNotFoundError(target_file)
    return torch.load(target_file, map_location="cpu", weights_only=False)


def build_model(config, state, device, precision=None) -> NCSNv2Deepest:
    """NCSNv2Deepest(config).cuda(); load_state_dict; eval   (test_score.py:59-63)."""
    diffuser = NCSNv2Deepest(config, precision=precision)
    diffuser = diffuser.to(device)
    diffuser.load_state_dict(state)
    diffuser.eval()
    return diffuser


def pick_device(gpu: int) -> Tuple[torch.device, int, int]:
    rank, ws, lr = sdist.init_from_env()
    if not torch.cuda.is_available():
        raise RuntimeError("a CUDA device is required (there is no CPU path)")
    dev = torch.device("cuda", lr if ws > 1 else gpu)
    torch.cuda.set_device(dev)
    return dev, rank, ws


def common_seed(seed: Optional[int], dev: torch.device, ws: int) -> int:
    """Seed of the sampler's Philox stream and of the measurement-noise generator: ``--seed`` when given (the
    reference seeds nothing), else a fresh draw -- rank 0's, so that every rank of a sharded run uses the same one."""
    s = int(seed) if seed is not None else int(np.random.randint(0, 2 ** 31 - 1))
    if ws > 1:
        t = torch.tensor([s], dtype=torch.int64, device=dev)
        torch.distributed.broadcast(t, src=0)
        s = int(t.item())
    return s


def finalize() -> None:
    """Tear down torch.distributed at the end of an entry point started under torchrun."""
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def ald_over_snr(diffuser, val_P, val_H, init_val_H, noise_range, alpha_step, beta_noise, sigma_end, num_levels,
                 steps_each, seed: int, id_base: int = 0, generator: Optional[torch.Generator] = None):
    """The SNR loop + ALD loop of test_score.py:118-171 for one (pilots, channels, init) set, with every SNR
    point stacked on the batch axis (exact: all ops are per-sample) and sharded over the ranks.

    val_P [B,Np,Nt], val_H / init_val_H [B,Nt,Nr] on the device; alpha_step / beta_noise scalars.
    Returns nmse [n_snr, steps, B] (float64 numpy, on every rank)."""
    B = val_H.shape[0]
    n_snr = len(noise_range)
    dev = val_H.device
    Ps, Hs, Xs, Ys, nvs = [], [], [], [], []
    for local_noise in noise_range:
        # val_Y = P @ H + sqrt(local_noise) * randn_like   (test_score.py:122-124)
        val_Y = torch.matmul(val_P, val_H)
        eps = torch.randn(val_Y.shape, dtype=val_Y.dtype, device=dev, generator=generator) \
            if generator is not None else torch.randn_like(val_Y)
        val_Y = val_Y + float(np.sqrt(local_noise)) * eps
        Ps.append(val_P); Hs.append(val_H); Xs.append(init_val_H.clone()); Ys.append(val_Y)
        nvs.append(torch.full((B,), float(local_noise), dtype=torch.float32, device=dev))
    P, H, X, Y, nv = (torch.cat(t, dim=0) for t in (Ps, Hs, Xs, Ys, nvs))
    total = n_snr * B

    def run(lo, hi):
        ids = torch.arange(id_base + lo, id_base + hi, dtype=torch.int64, device=dev)
        _, nlog = sampler.ald_run(diffuser, P[lo:hi], Y[lo:hi], X[lo:hi], H[lo:hi], noise_var=nv[lo:hi],
                                  alpha_step=alpha_step, beta=beta_noise, sigma_end=sigma_end, level_begin=0,
                                  level_end=num_levels, steps_each=steps_each, seed=seed, sample_ids=ids)
        return nlog

    full = sdist.run_sharded(run, total)                       # [steps, n_snr*B]
    steps = full.shape[0]
    return full.view(steps, n_snr, B).permute(1, 0, 2).double().cpu().numpy()


def maybe_plot(fn):
    try:
        from matplotlib import pyplot as plt  # noqa: F401
    except Exception:
        print("matplotlib not available: skipping the plot (results file is still written)")
        return
    fn()
