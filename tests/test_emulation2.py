"""CPU emulation of the engine-2 (tcgen05) plan: the C++ planner / packer of csrc/sbc2_plan.h is executed op by op on
its own byte layout (tests/emu/emu2.cpp replays every tcgen05.mma from its packed instruction record) and compared
with the golden vectors produced from the reference modules.  Catches planner, packing and addressing errors
without a GPU.  CPU only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from score_based_channels_b200 import params

from conftest import GOLDEN, REPO

EMU_DIR = os.path.join(REPO, "tests", "emu")


class Entry(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("shape", C.c_void_p), ("ndim", C.c_int32)]


class TensorInfo(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("fmt", C.c_int32), ("level", C.c_int32), ("C", C.c_int32),
                ("off", C.c_int64), ("bytes", C.c_int64), ("born", C.c_int32), ("died", C.c_int32)]


def load_emu2():
    so = os.path.join(EMU_DIR, "libemu2.so")
    csrc = os.path.join(REPO, "score_based_channels_b200", "csrc")
    srcs = [os.path.join(EMU_DIR, "emu2.cpp"), os.path.join(csrc, "sbc2_plan.h")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, srcs[0]])
    lib = C.CDLL(so)
    lib.emu2_create.restype = C.c_void_p
    lib.emu2_create.argtypes = [C.POINTER(Entry), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
    lib.emu2_free.argtypes = [C.c_void_p]
    lib.emu2_n_ops.argtypes = [C.c_void_p]
    lib.emu2_op_name.restype = C.c_char_p
    lib.emu2_op_name.argtypes = [C.c_void_p, C.c_int]
    lib.emu2_conv_flops.restype = C.c_longlong
    lib.emu2_conv_flops.argtypes = [C.c_void_p]
    lib.emu2_max_seg.argtypes = [C.c_void_p]
    lib.emu2_max_stage.argtypes = [C.c_void_p]
    lib.emu2_plan.restype = C.c_longlong
    lib.emu2_plan.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.emu2_forward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    return lib


def state_entries(sd):
    keep, ents = [], (Entry * len(sd))()
    for i, (k, v) in enumerate(sd.items()):
        a = np.ascontiguousarray(v, dtype=np.float32)
        shp = np.asarray(a.shape if a.ndim else (1,), dtype=np.int64)
        keep += [a, shp]
        ents[i] = Entry(k.encode(), a.ctypes.data, shp.ctypes.data, len(shp))
    return ents, keep


class Emu2:
    def __init__(self, lib, sd, ngf, H, W, stage_cap=32 * 1024):
        self.lib, self.H, self.W = lib, H, W
        ents, self._keep = state_entries(sd)
        err = C.create_string_buffer(256)
        self.h = lib.emu2_create(ents, len(sd), ngf, H, W, 2, stage_cap, err, 256)
        assert self.h, err.value

    def forward(self, x, reuse=True):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty_like(x)
        assert self.lib.emu2_forward(self.h, x.shape[0], int(reuse), x.ctypes.data, out.ctypes.data, None, -1) == 0
        return out

    def plan(self, S, reuse=True):
        n = C.c_int()
        info = (TensorInfo * 1024)()
        geo = np.zeros((4, 14), np.int32)
        ab = self.lib.emu2_plan(self.h, S, int(reuse), info, 1024, C.byref(n), geo.ctypes.data, None)
        return ab, list(info)[:n.value], geo

    def __del__(self):
        try:
            self.lib.emu2_free(self.h)
        except Exception:
            pass


@pytest.fixture(scope="module")
def emu2():
    return load_emu2()


@pytest.mark.parametrize("name,H,W", [("forward_ngf8.npz", 64, 16), ("forward_ngf8_32x8.npz", 32, 8),
                                      ("forward_ngf16.npz", 64, 16)])
def test_emulated_plan_matches_reference_golden(emu2, name, H, W):
    g = np.load(os.path.join(GOLDEN, name))
    ngf = int(g["ngf"])
    sd = params.random_state(ngf, seed=int(g["wseed"]))
    e = Emu2(emu2, sd, ngf, H, W)
    x, y, ref = g["x"], g["y"], g["out"]
    out = e.forward(x) / sd["sigmas"][y][:, None, None, None]
    for b in range(x.shape[0]):
        rel = np.linalg.norm(out[b] - ref[b]) / np.linalg.norm(ref[b])
        assert rel < 2e-5, (name, b, rel)       # fp16 hi/lo split operands: fp32-equivalent
    # group size must not matter: one sample alone gives the same numbers as inside the stacked group
    o1 = e.forward(x[1:2]) / sd["sigmas"][y[1:2]][:, None, None, None]
    np.testing.assert_allclose(o1[0], out[1], rtol=0, atol=1e-6 * np.abs(out[1]).max())


def test_plan_footprint_and_reuse(emu2):
    sd = params.random_state(8, seed=1)
    e = Emu2(emu2, sd, 8, 64, 16)
    assert emu2.emu2_conv_flops(e.h) == 51740672          # dense conv FLOP / forward / sample
    # shared-memory budget of two CTAs per SM: two weight buffers + staging ring + barriers + norm scratch
    smem = 2 * ((emu2.emu2_max_seg(e.h) + 127) // 128 * 128) + emu2.emu2_max_stage(e.h) + 24 * 8 + 4096
    assert smem <= 113 * 1024, smem
    a1, t1, _ = e.plan(1, True)
    a0, _, _ = e.plan(1, False)
    assert a1 < a0 / 4                                     # liveness reuse
    assert a1 < 600 * 1024                                 # per-sample arena stays L2 friendly (296 groups << 126 MB)
    # tensors that are live at the same time never overlap
    live = [t for t in t1 if t.fmt != 2]
    for i, a in enumerate(live):
        for b in live[i + 1:]:
            if a.born < b.died and b.born < a.died:
                assert a.off + a.bytes <= b.off or b.off + b.bytes <= a.off, (a.name, b.name)


@pytest.mark.parametrize("name", ["forward_deeper_ngf8.npz", "forward_ncsnv2_ngf8.npz"])
def test_emulated_plan_of_the_other_score_nets_matches_reference_golden(emu2, name):
    """NCSNv2Deeper / NCSNv2 (reference ncsnv2/models/ncsnv2.py:11-195): the C++ planner reads the architecture off the
    state-dict keys; golden vectors come from the reference modules (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLDEN, name))
    arch = str(g["arch"])
    sd = params.random_state(int(g["ngf"]), seed=int(g["wseed"]), arch=arch)
    e = Emu2(emu2, sd, int(g["ngf"]), 64, 16)
    x, y, ref = g["x"], g["y"], g["out"]
    out = e.forward(x) / sd["sigmas"][y][:, None, None, None]
    for b in range(x.shape[0]):
        rel = np.linalg.norm(out[b] - ref[b]) / np.linalg.norm(ref[b])
        assert rel < 2e-5, (name, b, rel)


def test_planner_rejects_malformed_state_dicts(emu2):
    """The library builds a model from caller-supplied tensors: shapes are validated before anything is indexed."""
    sd = params.random_state(8, seed=1)
    err = C.create_string_buffer(256)

    def create(d, ngf=8, H=64, W=16):
        ents, keep = state_entries(d)
        h = emu2.emu2_create(ents, len(d), ngf, H, W, 2, 32 * 1024, err, 256)
        if h:
            emu2.emu2_free(h)
        return bool(h), err.value.decode()

    assert create(sd)[0]
    bad = dict(sd); del bad["res3.1.conv2.weight"]
    ok, msg = create(bad); assert not ok and "res3.1.conv2.weight" in msg
    bad = dict(sd); bad["begin_conv.weight"] = sd["begin_conv.weight"].reshape(8, -1)          # not 4-D
    ok, msg = create(bad); assert not ok and "begin_conv.weight" in msg
    bad = dict(sd); bad["res1.0.normalize1.alpha"] = np.zeros(3, np.float32)
    ok, msg = create(bad); assert not ok and "normalize1" in msg
    bad = dict(sd); bad["end_conv.bias"] = np.zeros(5, np.float32)
    ok, msg = create(bad); assert not ok and "end_conv" in msg
    assert not create(sd, H=60)[0] and not create(sd, ngf=12)[0]
