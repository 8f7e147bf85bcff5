"""Packed-model handle: host packing (program.py) + the C-ABI model object (include/sbc.h)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _lib, program


ENGINE2_PRECISIONS = ("fp16x2",)      # tcgen05 engine: fp16 hi/lo split operands, fp32 accumulate (fp32-equivalent)
ALL_PRECISIONS = ENGINE2_PRECISIONS + tuple(program.PRECISIONS)


def engine1_runs_two_ctas_per_sm(state: Dict[str, np.ndarray], ngf: int, Nt: int, Nr: int, channels: int = 2) -> bool:
    """True when engine 1's planner (the C++ twin inside the library, host only) picks the two-CTAs-per-SM park plan for
    this shape -- the regime where engine 1 is the faster engine (the automatic engine choice of NCSNv2Deepest)."""
    L = _lib.lib()
    keep, ents = [], (_lib.StateEntry * len(state))()
    for i, (k, v) in enumerate(state.items()):
        a = np.ascontiguousarray(v, dtype=np.float32)
        shp = np.asarray(a.shape if a.ndim else (1,), dtype=np.int64)
        keep += [a, shp]
        ents[i] = _lib.StateEntry(k.encode(), a.ctypes.data, shp.ctypes.data, len(shp))
    h, v = C.c_void_p(), _lib.Plan1View()
    rc = L.sbc_plan1_build(ents, len(state), ngf, Nt, Nr, channels, int(L.sbc_threads_per_cta()), _lib.PREC_CODE["tf32x3"], -1,
                           C.byref(h), C.byref(v))
    if rc != 0:          # shapes engine 1 cannot plan at all (its planner says why): engine 2 decides
        return False
    geo = np.ctypeslib.as_array(C.cast(v.geo_table, C.POINTER(C.c_int32)), (v.n_geo, 8))
    misc = (64 + 2 * int(sum(g[5] - g[0] * g[1] for g in geo)) + 15) // 16 * 16
    two = v.park_floats > 0 and 4 * v.arena_floats + misc <= program.smem_budget_bytes(2)
    L.sbc_plan1_free(h)
    return bool(two)


class PackedModel:
    """Owns one ``sbc_model_create`` handle for a (state_dict, ngf, Nt, Nr) on one CUDA device."""

    def __init__(self, state: Dict[str, np.ndarray], ngf: int, Nt: int, Nr: int, device: int = 0,
                 channels: int = 2, precision: str = "tf32x3", planner: str = "python"):
        """``planner``: engine 1 only -- "python" = program.py (keeps ``self.prog`` for the simulator / profiling tools),
        "library" = the C++ twin inside the library (``sbc_model_create_from_state_ex``; identical op table)."""
        self.precision = precision
        self.ngf, self.Nt, self.Nr, self.channels, self.device = ngf, Nt, Nr, channels, device
        self.sigmas = np.ascontiguousarray(state["sigmas"], dtype=np.float32)
        self.prog = None
        if precision in ENGINE2_PRECISIONS or planner == "library":
            # the library plans and packs from the plain state dict: engine 2 (tcgen05), or engine 1 through the C++ twin
            # of program.py
            keep, ents = [], (_lib.StateEntry * len(state))()
            for i, (k, v) in enumerate(state.items()):
                a = np.ascontiguousarray(v, dtype=np.float32)
                shp = np.asarray(a.shape if a.ndim else (1,), dtype=np.int64)
                keep += [a, shp]
                ents[i] = _lib.StateEntry(k.encode(), a.ctypes.data, shp.ctypes.data, len(shp))
            h = C.c_void_p()
            _lib.check(_lib.lib().sbc_model_create_from_state_ex(ents, len(state), ngf, Nt, Nr, channels, device,
                                                                 _lib.PREC_CODE[precision], C.byref(h)),
                       "sbc_model_create_from_state_ex")
            self.handle = h
            self.conv_flops = int(self.info().conv_flops_per_forward)
            return
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
                           self.sigmas.ctypes.data, self.sigmas.size, p.conv_flops, p.nthreads, p.park_floats)
        h = C.c_void_p()
        _lib.check(_lib.lib().sbc_model_create(C.byref(d), device, C.byref(h)), "sbc_model_create")
        self.handle = h
        self.conv_flops = int(p.conv_flops)

    def info(self) -> _lib.Info:
        out = _lib.Info()
        _lib.check(_lib.lib().sbc_query(self.handle, C.byref(out)), "sbc_query")
        return out

    # ---- engine-2 debug views (tests / tools) ----
    def debug_plan(self, S: int, reuse: bool = True):
        """(arena_bytes, [TensorInfo], geo[4][12]) of the engine-2 plan for group size S."""
        n, ab = C.c_int32(), C.c_int64()
        info = (_lib.TensorInfo * 2048)()
        geo = np.zeros((4, 14), np.int32)
        _lib.check(_lib.lib().sbc_debug_plan(self.handle, S, int(reuse), info, 2048, C.byref(n), C.byref(ab),
                                             geo.ctypes.data), "sbc_debug_plan")
        return int(ab.value), list(info)[:n.value], geo

    def op_names(self):
        L = _lib.lib()
        return [L.sbc_op_name(self.handle, i).decode() for i in range(self.info().n_ops)]

    def close(self) -> None:
        if getattr(self, "handle", None):
            _lib.lib().sbc_model_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
