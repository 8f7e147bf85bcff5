"""CPU thread-emulation of the kernel's per-thread op bodies (csrc/sbc_ops.h, shared verbatim with
the CUDA kernel) against the schedule simulator and the reference goldens.  Catches tiling /
indexing errors in the device code without a GPU.  CPU only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from score_based_channels_b200 import params, program

from conftest import GOLDEN, REPO

EMU_DIR = os.path.join(REPO, "tests", "emu")


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(EMU_DIR, "libemu.so")
    csrc = os.path.join(REPO, "score_based_channels_b200", "csrc")
    srcs = [os.path.join(EMU_DIR, "emu.cpp")] + [os.path.join(csrc, f) for f in ("sbc_ops.h", "sbc_mma.h", "sbc_program.h")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, srcs[0]])
    lib = C.CDLL(so)
    lib.emu_run_program.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.emu_langevin_step.restype = C.c_float
    lib.emu_langevin_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint64, C.c_uint64,
                                      C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.emu_noise.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p]
    return lib


def run_emu(lib, prog, x, stop=-1):
    # garbage-filled arena: every op must (re)establish the zero halos it relies on
    arena = np.random.default_rng(5).standard_normal(prog.arena_floats).astype(np.float32) * 100
    park = np.random.default_rng(6).standard_normal(max(prog.park_floats, 4)).astype(np.float32) * 100
    prog.write_input(arena, np.asarray(x, np.float32))
    tab = np.ascontiguousarray(prog.op_table())
    geo = np.ascontiguousarray(prog.geo_table())
    rc = lib.emu_run_program(tab.ctypes.data, tab.shape[0], geo.ctypes.data, prog.blob.ctypes.data, arena.ctypes.data,
                             park.ctypes.data, prog.nthreads, stop)
    assert rc == 0
    run_emu.park = park
    return arena


# tolerance on the per-sample relative L2 error of one forward, per precision mode
FWD_TOL = {"tf32x3": 2e-5, "tf32": 6e-3}


@pytest.mark.parametrize("prec", ["tf32x3", "tf32"])
@pytest.mark.parametrize("name,H,W", [("forward_ngf8.npz", 64, 16), ("forward_ngf8_32x8.npz", 32, 8),
                                      ("forward_ngf16.npz", 64, 16)])
def test_emulated_forward_matches_reference_golden(emu, name, H, W, prec):
    g = np.load(os.path.join(GOLDEN, name))
    sd = params.random_state(int(g["ngf"]), seed=int(g["wseed"]))
    prog = program.build_program(sd, int(g["ngf"]), H, W, precision=prec)
    sig = sd["sigmas"]
    for b in range(g["x"].shape[0]):
        arena = run_emu(emu, prog, g["x"][b])
        out = prog.read_output(arena) / sig[int(g["y"][b])]
        rel = np.linalg.norm(out - g["out"][b]) / np.linalg.norm(g["out"][b])
        assert rel < FWD_TOL[prec], (name, b, rel)


@pytest.mark.parametrize("prec", ["tf32x3"])
def test_emulated_ops_match_simulator_op_by_op(emu, prec):
    """Prefix runs: the arena after k ops, emulated device code vs the torch simulator (a failure
    names the op)."""
    sd = params.random_state(8, seed=1)
    prog = program.build_program(sd, 8, 64, 16, precision=prec)
    x = (np.random.default_rng(0).standard_normal((2, 64, 16)) * 3).astype(np.float32)
    assert prog.park_floats > 0, "the 64x16 ngf-8 plan is expected to be the two-CTAs-per-SM (park) plan"
    for k in list(range(1, 40)) + list(range(40, len(prog.ops), 7)) + list(range(len(prog.ops) - 35, len(prog.ops) + 1)):
        _, ra = program.simulate(prog, torch.from_numpy(x), upto=k)
        ea = run_emu(emu, prog, x, stop=k)
        op = prog.ops[k - 1]
        if op.flags & program.F_COMPACT:
            e, r = prog.read_output(ea), prog.read_output(ra.numpy())
            assert np.abs(e - r).max() / (np.abs(r).max() + 1e-6) < 5e-5, (k - 1, op.name)
            continue
        if op.kind == program.OP_SPILL or (op.kind == program.OP_FILL and op.cin == 0):
            n = 4 * op.MT   # whole-tensor copies into the park area / raw sampler state coming back
            if op.kind == program.OP_SPILL:
                e, r = run_emu.park[op.dst:op.dst + n], ra.park.numpy()[op.dst:op.dst + n]
            else:
                e, r = ea[op.dst:op.dst + n], ra.numpy()[op.dst:op.dst + n]
            assert np.abs(e - r).max() / (np.abs(r).max() + 1e-6) < 5e-5, (k - 1, op.name)
            continue
        acc_g = bool(op.flags & program.F_ACC_G)
        for fi, off in enumerate((op.dst, op.acc, op.edst)):
            if off >= 0:
                eb, rb = (run_emu.park, ra.park.numpy()) if (fi == 1 and acc_g) else (ea, ra.numpy())
                e = prog.read(eb, off, op.cout, op.oh, op.ow)
                r = prog.read(rb, off, op.cout, op.oh, op.ow)
                d, sc = np.abs(e - r).max(), np.abs(r).max() + 1e-6
                assert d / sc < 5e-5, (k - 1, op.name, op.kind, d, sc)


def test_emulated_langevin_step_matches_oracle(emu):
    g = np.load(os.path.join(GOLDEN, "ald_cfg1.npz"))
    sd = params.random_state(8, seed=int(g["wseed"]))
    prog = program.build_program(sd, 8, 64, 16)
    sig = sd["sigmas"]
    b, Nt, Nr, Np = 1, 64, 16, g["P"].shape[1]
    lvl = int(g["levels"][0])
    sigma = float(sig[lvl])
    alpha = float(g["alpha_step"]) * (sigma / float(g["sigma_end"])) ** 2
    den = float(g["noise_var"]) / 2. + sigma ** 2
    nscale = np.sqrt(2 * alpha * float(g["beta"]))
    X0 = g["X0"][b]
    xr = np.stack([X0.real, X0.imag]).astype(np.float32)
    arena = run_emu(emu, prog, xr)
    P = np.ascontiguousarray(g["P"][b]); Y = np.ascontiguousarray(g["Y"][b]); H = np.ascontiguousarray(g["H"][b])
    en = np.ascontiguousarray(g["ext_noise"][0, b])
    tot = emu.emu_langevin_step(arena.ctypes.data, prog.in_off, prog.out_off, prog.post_off, P.ctypes.data,
                                Y.ctypes.data, H.ctypes.data, en.ctypes.data, sigma, alpha, den, nscale, 0, 0, 0, Nt,
                                Nr, Np, prog.nthreads)
    xv = arena[prog.in_off:prog.in_off + 2 * Nt * Nr].reshape(Nt, Nr, 2)
    x1 = xv[..., 0] + 1j * xv[..., 1]
    ref = g["xs"][0, b]
    assert np.abs(x1 - ref).max() < 1e-5 * np.abs(ref).max()
    nm = tot / np.sum(np.abs(H) ** 2)
    assert abs(nm - g["nmse"][0, b]) < 2e-5 * g["nmse"][0, b]


def test_device_noise_matches_oracle_noise(emu):
    out = np.empty(1024, np.complex64)
    emu.emu_noise(77, 5, 1234, 1024, out.ctypes.data)
    ref = orc.noise(77, 5, 1234, 1024)
    assert np.allclose(out, ref, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("H,W", [(24, 8), (40, 24), (72, 16)])
def test_emulated_forward_on_non_power_of_two_geometries(emu, H, W):
    """Row widths / pixel counts that are not powers of two take the division fall-backs of the device
    code (sbc_div, sbc_for_pixels, direct max-pool); checked against the oracle."""
    sd = params.random_state(8, seed=5)
    x = (np.random.default_rng(1).standard_normal((2, H, W)) * 2).astype(np.float32)
    ref = orc.OracleNet(sd, 8, H, W).forward(x[None], np.array([1200]))[0]
    prog = program.build_program(sd, 8, H, W)
    out = prog.read_output(run_emu(emu, prog, x)) / sd["sigmas"][1200]
    assert np.linalg.norm(out - ref) / np.linalg.norm(ref) < 2e-5
