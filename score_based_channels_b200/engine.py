"""Packed-model handle: host packing (program.py) + the C-ABI model object (include/sbc.h)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _lib, program


class PackedModel:
    """Owns one ``sbc_model_create`` handle for a (state_dict, ngf, Nt, Nr) on one CUDA device."""

    def __init__(self, state: Dict[str, np.ndarray], ngf: int, Nt: int, Nr: int, device: int = 0,
                 channels: int = 2, precision: str = "tf32x3"):
        self.precision = precision
        nthreads = int(_lib.lib().sbc_threads_per_cta())
        self.prog = program.build_program(state, ngf, Nt, Nr, channels, nthreads=nthreads, precision=precision)
        self.sigmas = np.ascontiguousarray(state["sigmas"], dtype=np.float32)
        self.ngf, self.Nt, self.Nr, self.channels, self.device = ngf, Nt, Nr, channels, device
        self._tab = np.ascontiguousarray(self.prog.op_table())
        self._geo = np.ascontiguousarray(self.prog.geo_table())
        p = self.prog
        d = _lib.ModelDesc(ngf, Nt, Nr, channels, self._tab.ctypes.data, self._tab.shape[0], self._geo.ctypes.data,
                           len(p.geos), p.blob.ctypes.data,
                           p.blob.size, p.arena_floats, p.in_off, p.out_off, p.post_off, p.max_w_len,
                           self.sigmas.ctypes.data, self.sigmas.size, p.conv_flops, p.nthreads)
        h = C.c_void_p()
        _lib.check(_lib.lib().sbc_model_create(C.byref(d), device, C.byref(h)), "sbc_model_create")
        self.handle = h

    def info(self) -> _lib.Info:
        out = _lib.Info()
        _lib.check(_lib.lib().sbc_query(self.handle, C.byref(out)), "sbc_query")
        return out

    def close(self) -> None:
        if getattr(self, "handle", None):
            _lib.lib().sbc_model_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
