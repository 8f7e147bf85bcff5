"""ctypes front-end of the C oracle (``oracle/sbc_oracle.c``) -- TEST INFRASTRUCTURE ONLY.

The oracle restates, on the CPU in fp32, ``NCSNv2Deepest.forward``
(reference ``ncsnv2/models/ncsnv2.py:269-300``) and the annealed-Langevin loop body of
``test_score.py:135-171``.  It is pinned by ``tests/test_oracle.py`` against golden vectors
produced from the reference's own Python modules (``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libsbc_oracle.so")
    src = os.path.join(_HERE, "sbc_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libsbc_oracle.so"])
    return so


class _AldArgs(C.Structure):
    _fields_ = [("B", C.c_int), ("Nt", C.c_int), ("Nr", C.c_int), ("Np", C.c_int),
                ("level_begin", C.c_int), ("level_end", C.c_int), ("steps_each", C.c_int),
                ("P", C.c_void_p), ("Y", C.c_void_p), ("X", C.c_void_p), ("H_oracle", C.c_void_p),
                ("noise_var", C.c_void_p), ("alpha_step", C.c_void_p), ("beta", C.c_void_p),
                ("sigmas", C.c_void_p), ("sigma_end", C.c_double), ("nmse_log", C.c_void_p),
                ("seed", C.c_uint64), ("sample_ids", C.c_void_p), ("ext_noise", C.c_void_p),
                ("dc_boost", C.c_void_p), ("stop_step", C.c_void_p)]


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_net_create.restype = C.c_void_p
        _LIB.orc_net_create.argtypes = [C.c_int] * 4
        _LIB.orc_net_set.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_long]
        _LIB.orc_net_free.argtypes = [C.c_void_p]
        _LIB.orc_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _LIB.orc_ald_run.argtypes = [C.c_void_p, C.POINTER(_AldArgs)]
        _LIB.orc_noise.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p]
        _LIB.orc_set_num_threads.argtypes = [C.c_int]
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.complex64)


class OracleNet:
    """fp32 CPU NCSNv2Deepest built from a reference-keyed state dict (numpy arrays)."""

    def __init__(self, state: Dict[str, np.ndarray], ngf: int, H: int, W: int, channels: int = 2):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_net_create(ngf, H, W, channels))
        self.H, self.W, self.channels = H, W, channels
        self.sigmas = _f32(state["sigmas"])
        for k, v in state.items():
            a = _f32(v)
            self.L.orc_net_set(self.h, k.encode(), a.ctypes.data, a.size)

    def __del__(self):
        try:
            self.L.orc_net_free(self.h)
        except Exception:
            pass

    def forward(self, x: np.ndarray, y: np.ndarray) -> np.ndarray:
        """x [B,channels,H,W] f32, y [B] int labels -> score [B,channels,H,W] (ncsnv2.py:269-300)."""
        x = _f32(x)
        sig = _f32(self.sigmas[np.asarray(y, dtype=np.int64)])
        out = np.empty_like(x)
        self.L.orc_forward(self.h, x.ctypes.data, sig.ctypes.data, out.ctypes.data, x.shape[0])
        return out

    def dsm_losses(self, samples: np.ndarray, labels: np.ndarray, z: np.ndarray, anneal_power: float = 2.0) -> np.ndarray:
        """Per-sample anneal_dsm_score_estimation terms (reference ncsnv2/losses/dsm.py:13-31, before the mean) for the
        randn draw ``z``: fp32, same operation order as the reference."""
        x, z = _f32(samples), _f32(z)
        B = x.shape[0]
        sig = _f32(self.sigmas[np.asarray(labels, dtype=np.int64)]).reshape(B, 1, 1, 1)
        noise = z * sig
        target = np.float32(-1.0) / (sig ** 2) * noise
        scores = self.forward(x + noise, labels)
        d = (scores.reshape(B, -1) - target.reshape(B, -1)).astype(np.float32)
        return (np.float32(0.5) * (d ** 2).sum(axis=-1, dtype=np.float32) * sig.reshape(B) ** np.float32(anneal_power)).astype(np.float32)

    def ald(self, P, Y, X0, H=None, *, noise_var, alpha_step, beta, sigma_end, level_begin=0,
            level_end=None, steps_each=3, seed=0, sample_ids=None, ext_noise=None, log=True, dc_boost=None,
            stop_step=None):
        """Annealed Langevin loop of test_score.py:135-171 for a batch with per-sample scalars.

        P [B,Np,Nt], Y [B,Np,Nr], X0/H [B,Nt,Nr] complex64.  Returns (X_final, nmse_log[steps,B] or None)."""
        P, Y, X = _c64(P), _c64(Y), _c64(X0).copy()
        B, Np, Nt = P.shape
        Nr = Y.shape[2]
        assert (Nt, Nr) == (self.H, self.W)
        if level_end is None:
            level_end = self.sigmas.size
        nsteps = (level_end - level_begin) * steps_each
        Hc = _c64(H) if H is not None else None
        nv, al, be = (_f32(np.broadcast_to(np.asarray(v, np.float32), (B,))) for v in (noise_var, alpha_step, beta))
        nlog = np.zeros((nsteps, B), np.float32) if (log and Hc is not None) else None
        ids = np.ascontiguousarray(sample_ids, dtype=np.uint64) if sample_ids is not None else None
        en = _c64(ext_noise) if ext_noise is not None else None
        if en is not None:
            assert en.shape == (nsteps, B, Nt, Nr)
        db = _f32(np.broadcast_to(np.asarray(dc_boost, np.float32), (B,))) if dc_boost is not None else None
        st = np.ascontiguousarray(np.broadcast_to(np.asarray(stop_step, np.int32), (B,))) if stop_step is not None else None
        a = _AldArgs(B, Nt, Nr, Np, level_begin, level_end, steps_each, P.ctypes.data, Y.ctypes.data,
                     X.ctypes.data, Hc.ctypes.data if Hc is not None else None, nv.ctypes.data, al.ctypes.data,
                     be.ctypes.data, self.sigmas.ctypes.data, float(sigma_end),
                     nlog.ctypes.data if nlog is not None else None, int(seed),
                     ids.ctypes.data if ids is not None else None, en.ctypes.data if en is not None else None,
                     db.ctypes.data if db is not None else None, st.ctypes.data if st is not None else None)
        self.L.orc_ald_run(self.h, C.byref(a))
        return X, nlog


def noise(seed: int, sid: int, step: int, n_elem: int) -> np.ndarray:
    """CN(0,1) Philox noise of the project's RNG contract (include/sbc.h) as complex64 [n_elem]."""
    out = np.empty(n_elem, np.complex64)
    lib().orc_noise(int(seed), int(sid), int(step), int(n_elem), out.ctypes.data)
    return out


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


def num_threads() -> int:
    return int(lib().orc_num_threads())
