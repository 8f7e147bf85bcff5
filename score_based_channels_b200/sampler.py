"""Annealed-Langevin sampler: torch-tensor front end of ``sbc_ald_run`` (include/sbc.h).

Replaces the loop body of reference ``test_score.py:135-171`` / ``tune_hparams_score.py:112-148``:
for every sigma level in [level_begin, level_end) and every inner step the fused kernel evaluates
the score network, the data-consistency gradient ``P^H (P x - y)``, the Langevin update with Philox
noise and the per-step NMSE -- one launch for the whole range, no host round trips."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from .ncsnv2 import NCSNv2Deepest


def _per_sample(v, B, device):
    t = torch.as_tensor(v, dtype=torch.float32, device=device)
    if t.dim() == 0:
        t = t.expand(B)
    if t.numel() != B:
        raise ValueError("per-sample array must have %d entries" % B)
    return t.contiguous()


@torch.no_grad()
def ald_run(model: NCSNv2Deepest, P: torch.Tensor, Y: torch.Tensor, X0: torch.Tensor,
            H: Optional[torch.Tensor] = None, *, noise_var, alpha_step, beta, sigma_end: Optional[float] = None,
            level_begin: int = 0, level_end: Optional[int] = None, steps_each: int = 3, seed: int = 0,
            sample_ids: Optional[torch.Tensor] = None, ext_noise: Optional[torch.Tensor] = None,
            log_nmse: bool = True, inplace: bool = False, dc_boost=None,
            stop_step: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Run ALD for a batch of independent channel realisations.

    P [B,Np,Nt], Y [B,Np,Nr], X0 / H [B,Nt,Nr]: complex64 CUDA tensors (``forward``, ``y``,
    ``current``, ``oracle`` of test_score.py:126-131).  ``noise_var`` (= local_noise,
    test_score.py:75), ``alpha_step``, ``beta`` are scalars or per-sample arrays.  ``dc_boost`` (scalar or
    per-sample, reference test_mmse.py:25,246) multiplies the data-consistency term; ``stop_step`` (int or
    per-sample int32, test_mmse.py:173,260-263) is the index of the last step a sample executes -- NMSE-log rows
    after it are NaN.

    Returns (X_final [B,Nt,Nr] complex64, nmse_log [steps,B] fp32 or None)."""
    if not (P.is_cuda and Y.is_cuda and X0.is_cuda):
        raise RuntimeError("ald_run needs CUDA tensors; there is no CPU path")
    dev = X0.device
    B, Np, Nt = P.shape
    Nr = Y.shape[2]
    if X0.shape != (B, Nt, Nr) or Y.shape != (B, Np, Nr):
        raise ValueError("inconsistent shapes P %s Y %s X %s" % (tuple(P.shape), tuple(Y.shape), tuple(X0.shape)))
    # torch's lazy conj / neg bits do not change data_ptr(): materialise them before handing raw pointers over
    _mat = lambda t: t.to(torch.complex64).resolve_conj().resolve_neg().contiguous()
    P = _mat(P)
    Y = _mat(Y)
    X = _mat(X0)
    if not inplace and X.data_ptr() == X0.data_ptr():
        X = X.clone()
    Hc = _mat(H) if H is not None else None
    if level_end is None:
        level_end = model.num_classes
    if sigma_end is None:
        sigma_end = float(model.config.model.sigma_end)
    nsteps = (level_end - level_begin) * steps_each
    nv, al, be = (_per_sample(v, B, dev) for v in (noise_var, alpha_step, beta))
    nlog = None
    if log_nmse and Hc is not None:
        if stop_step is not None:      # rows after a sample's stop stay NaN
            nlog = torch.full((nsteps, B), float("nan"), dtype=torch.float32, device=dev)
        else:
            nlog = torch.empty((nsteps, B), dtype=torch.float32, device=dev)
    db = _per_sample(dc_boost, B, dev) if dc_boost is not None else None
    st = None
    if stop_step is not None:
        st = torch.as_tensor(stop_step, dtype=torch.int32, device=dev)
        st = (st.expand(B) if st.dim() == 0 else st).contiguous()
        if st.numel() != B:
            raise ValueError("stop_step must have %d entries" % B)
    ids = sample_ids.to(device=dev, dtype=torch.int64).contiguous() if sample_ids is not None else None
    en = None
    if ext_noise is not None:
        en = _mat(ext_noise.to(device=dev))
        if en.shape != (nsteps, B, Nt, Nr):
            raise ValueError("ext_noise must be [steps,B,Nt,Nr]")
    pm = model.packed(Nt, Nr, dev)
    a = _lib.AldArgs(B, Nt, Nr, Np, level_begin, level_end, steps_each, P.data_ptr(), Y.data_ptr(), X.data_ptr(),
                     Hc.data_ptr() if Hc is not None else None, nv.data_ptr(), al.data_ptr(), be.data_ptr(),
                     float(sigma_end), nlog.data_ptr() if nlog is not None else None, int(seed) & (2 ** 64 - 1),
                     ids.data_ptr() if ids is not None else None, en.data_ptr() if en is not None else None,
                     db.data_ptr() if db is not None else None, st.data_ptr() if st is not None else None)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().sbc_ald_run(pm.handle, C.byref(a), C.c_void_p(stream)), "sbc_ald_run")
    return X, nlog
