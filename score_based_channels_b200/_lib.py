"""ctypes binding of the C ABI in ``include/sbc.h`` (``libsbc_b200.so``, built by
``__graft_entry__.build()``).  There is deliberately no fallback: if the CUDA library is missing
or a call fails, a RuntimeError is raised."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SBC_LIB") or os.path.join(_HERE, "libsbc_b200.so")   # SBC_LIB: experiment builds (tools/)

EXPORTS = ("sbc_version", "sbc_threads_per_cta", "sbc_last_error", "sbc_model_create", "sbc_model_create_from_state",
           "sbc_model_free", "sbc_query", "sbc_forward", "sbc_ald_run", "sbc_forward_host", "sbc_ald_run_host",
           "sbc_debug_arena", "sbc_set_profile_buffer", "sbc_debug_plan", "sbc_debug_run", "sbc_op_name", "sbc_op_kind",
           "sbc_model_create_from_state_ex", "sbc_plan1_build", "sbc_plan1_free", "sbc_dsm_loss")
PREC_CODE = {"fp16x2": 0, "tf32x3": 1, "tf32": 2}


class ModelDesc(C.Structure):
    _fields_ = [("ngf", C.c_int32), ("Nt", C.c_int32), ("Nr", C.c_int32), ("channels", C.c_int32),
                ("op_table", C.c_void_p), ("n_ops", C.c_int32), ("geo_table", C.c_void_p), ("n_geo", C.c_int32),
                ("blob", C.c_void_p), ("blob_floats", C.c_int64),
                ("arena_floats", C.c_int32), ("in_off", C.c_int32), ("out_off", C.c_int32), ("post_off", C.c_int32),
                ("max_w_len", C.c_int32), ("sigmas", C.c_void_p), ("n_sigmas", C.c_int32), ("conv_flops", C.c_int64),
                ("nthreads", C.c_int32), ("park_floats", C.c_int32)]


class Info(C.Structure):
    _fields_ = [("version", C.c_int32), ("device", C.c_int32), ("num_sms", C.c_int32), ("threads_per_cta", C.c_int32),
                ("arena_in_smem", C.c_int32), ("weights_staged", C.c_int32), ("smem_bytes_per_cta", C.c_int64),
                ("arena_bytes", C.c_int64), ("conv_flops_per_forward", C.c_int64), ("kernel_launches", C.c_int64),
                ("engine", C.c_int32), ("ctas_per_sm", C.c_int32), ("group_size", C.c_int32), ("n_ops", C.c_int32)]


class StateEntry(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("shape", C.c_void_p), ("ndim", C.c_int32)]


class Plan1View(C.Structure):
    _fields_ = [("op_table", C.c_void_p), ("n_ops", C.c_int32), ("geo_table", C.c_void_p), ("n_geo", C.c_int32),
                ("blob", C.c_void_p), ("blob_floats", C.c_int64), ("arena_floats", C.c_int32), ("in_off", C.c_int32),
                ("out_off", C.c_int32), ("post_off", C.c_int32), ("max_w_len", C.c_int32), ("park_floats", C.c_int32),
                ("nthreads", C.c_int32), ("conv_flops", C.c_int64)]


class TensorInfo(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("fmt", C.c_int32), ("level", C.c_int32), ("C", C.c_int32),
                ("off", C.c_int64), ("bytes", C.c_int64), ("born", C.c_int32), ("died", C.c_int32)]


class AldArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("Nt", C.c_int32), ("Nr", C.c_int32), ("Np", C.c_int32),
                ("level_begin", C.c_int32), ("level_end", C.c_int32), ("steps_each", C.c_int32),
                ("P", C.c_void_p), ("Y", C.c_void_p), ("X", C.c_void_p), ("H_oracle", C.c_void_p),
                ("noise_var", C.c_void_p), ("alpha_step", C.c_void_p), ("beta", C.c_void_p),
                ("sigma_end", C.c_double), ("nmse_log", C.c_void_p), ("seed", C.c_uint64),
                ("sample_ids", C.c_void_p), ("ext_noise", C.c_void_p), ("dc_boost", C.c_void_p),
                ("stop_step", C.c_void_p)]


_lib = None


def lib():
    """Load the CUDA library; fail loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("CUDA extension %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.sbc_version.restype = C.c_int
        L.sbc_threads_per_cta.restype = C.c_int
        L.sbc_last_error.restype = C.c_char_p
        L.sbc_model_create.argtypes = [C.POINTER(ModelDesc), C.c_int, C.POINTER(C.c_void_p)]
        L.sbc_model_free.argtypes = [C.c_void_p]
        L.sbc_query.argtypes = [C.c_void_p, C.POINTER(Info)]
        L.sbc_forward.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.c_void_p, C.c_int32,
                                  C.c_void_p]
        L.sbc_dsm_loss.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int32,
                                   C.c_void_p]
        L.sbc_ald_run.argtypes = [C.c_void_p, C.POINTER(AldArgs), C.c_void_p]
        L.sbc_forward_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.sbc_ald_run_host.argtypes = [C.c_void_p, C.POINTER(AldArgs)]
        L.sbc_debug_arena.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.sbc_set_profile_buffer.argtypes = [C.c_void_p, C.c_void_p]
        L.sbc_model_create_from_state.argtypes = [C.POINTER(StateEntry), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                                  C.c_int32, C.c_int, C.POINTER(C.c_void_p)]
        L.sbc_model_create_from_state_ex.argtypes = [C.POINTER(StateEntry), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                                     C.c_int32, C.c_int, C.c_int32, C.POINTER(C.c_void_p)]
        L.sbc_plan1_build.argtypes = [C.POINTER(StateEntry), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(Plan1View)]
        L.sbc_plan1_free.argtypes = [C.c_void_p]
        L.sbc_debug_plan.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_void_p]
        L.sbc_debug_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.sbc_op_name.restype = C.c_char_p
        L.sbc_op_name.argtypes = [C.c_void_p, C.c_int32]
        L.sbc_op_kind.argtypes = [C.c_void_p, C.c_int32]
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (what, rc, lib().sbc_last_error().decode()))
