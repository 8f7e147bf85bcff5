"""Synthetic "CDL-shaped" MIMO channels, pilots and received pilots.

The reference's channels come from MATLAB's ``nrCDLChannel`` (``matlab/generate_data.m``), which
cannot run here; throughput does not depend on the channel values, and parity tests only need
inputs that both implementations share.  Shapes / normalisation follow ``loaders.py:30-55,88-91``
and ``test_score.py:108-124``:

* ``H``      [B, Nt, Nr] complex64 -- the (normalised) Hermitian channel the sampler estimates;
* ``P``      [B, Np, Nt] complex64 -- ``val_P = conj(P^T)`` of QPSK pilots ``(+-1 +-1j)/sqrt(2)``;
* ``Y``      [B, Np, Nr] complex64 -- ``P @ H + sqrt(noise_var) * CN(0,1)``, ``noise_var = 10^(-snr/10)*Nt``.
"""
from __future__ import annotations

import numpy as np


def cdl_like_channels(B: int, Nt: int = 64, Nr: int = 16, n_clusters: int = 24, seed: int = 4321) -> np.ndarray:
    """H[b] = sum_l g_l a_t(phi_l) a_r(theta_l)^H, half-wavelength ULAs, unit average entry power."""
    out = np.empty((B, Nt, Nr), np.complex64)
    kt, kr = np.arange(Nt)[:, None], np.arange(Nr)[:, None]
    pw = np.exp(-0.25 * np.arange(n_clusters))
    pw /= pw.sum()
    for b in range(B):
        rng = np.random.default_rng(seed + b)
        centre_t, centre_r = rng.uniform(-60, 60, 2)
        phi = np.deg2rad(np.clip(centre_t + rng.normal(0, 15, n_clusters), -89, 89))
        theta = np.deg2rad(np.clip(centre_r + rng.normal(0, 25, n_clusters), -89, 89))
        g = np.sqrt(pw / 2) * (rng.standard_normal(n_clusters) + 1j * rng.standard_normal(n_clusters))
        at = np.exp(1j * np.pi * kt * np.sin(phi)[None, :])          # [Nt, L]
        ar = np.exp(1j * np.pi * kr * np.sin(theta)[None, :])        # [Nr, L]
        out[b] = ((at * g[None, :]) @ ar.conj().T).astype(np.complex64)
    out /= np.sqrt(np.mean(np.abs(out) ** 2))
    return out


def qpsk_pilots(B: int, Nt: int, Np: int, seed: int = 1234) -> np.ndarray:
    """val_P [B, Np, Nt] = conj(transpose(pilots [B, Nt, Np])) with QPSK entries (loaders.py:52-55,
    test_score.py:108-110)."""
    rng = np.random.default_rng(seed)
    p = (2 * rng.integers(0, 2, (B, Nt, Np)) - 1 + 1j * (2 * rng.integers(0, 2, (B, Nt, Np)) - 1)) / np.sqrt(2)
    return np.ascontiguousarray(np.conj(np.transpose(p, (0, 2, 1))), dtype=np.complex64)


def cn01(shape, rng) -> np.ndarray:
    """Unit-power circular complex Gaussian (re, im ~ N(0, 1/2)) -- what torch.randn_like gives for complex."""
    return ((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2)).astype(np.complex64)


def received_pilots(P: np.ndarray, H: np.ndarray, noise_var, seed: int = 99) -> np.ndarray:
    """val_Y = P @ H + sqrt(noise_var) * CN(0,1)  (test_score.py:122-124); noise_var scalar or [B]."""
    rng = np.random.default_rng(seed)
    nv = np.asarray(noise_var, np.float64).reshape(-1, 1, 1)
    Y = np.matmul(P, H)
    return (Y + np.sqrt(nv) * cn01(Y.shape, rng)).astype(np.complex64)


def snr_to_noise_var(snr_db, Nt: int) -> np.ndarray:
    """local_noise = 10^(-snr/10) * Nt  (test_score.py:75)."""
    return (10.0 ** (-np.asarray(snr_db, np.float64) / 10.0) * Nt)
