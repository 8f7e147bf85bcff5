"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes, include/sbc.h), against
(i) the committed golden vectors produced by the reference's own modules and (ii) the CPU oracle on
the same seeded inputs; plus size-independent properties.  Tolerances (fp32 path): per-sample
relative L2 of a forward <= 2e-5; state after each ALD run <= 1e-5 * max|x| (SURVEY.md 8(d),
config 1); NMSE log relative <= 1e-4."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, REPO

pytestmark = pytest.mark.gpu

from score_based_channels_b200 import _lib, params, program, sampler, synth  # noqa: E402
from score_based_channels_b200.models import make_model  # noqa: E402

SIGMA_END = 2.599515446446343e-4


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


# precision mode -> (forward rel-L2 tol, ALD state tol relative to max|x|, NMSE-log rtol)
TOL = {"tf32x3": (2e-5, 1e-5, 1e-4), "fp16x2": (2e-5, 1e-5, 1e-4), "tf32": (6e-3, 2e-3, 5e-3)}
PRECS = list(TOL)


def _model(ngf, wseed, dev, prec="tf32x3"):
    sd = params.random_state(ngf, seed=wseed)
    return sd, make_model(sd, ngf=ngf, precision=prec).to(dev)


def _rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))


def test_native_library_is_the_one_running(dev):
    L = _lib.lib()
    assert L.sbc_version() == 210
    sd, m = _model(8, 1, dev)
    info = m.packed(64, 16, dev).info()
    assert info.num_sms >= 100 and info.threads_per_cta == L.sbc_threads_per_cta() and info.threads_per_cta in (256, 512, 1024)
    assert info.arena_in_smem == 1 and info.weights_staged == 1


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("name,H,W", [("forward_ngf8.npz", 64, 16), ("forward_ngf8_32x8.npz", 32, 8),
                                      ("forward_ngf16.npz", 64, 16)])
def test_forward_matches_reference_golden(dev, name, H, W, prec):
    g = np.load(os.path.join(GOLDEN, name))
    sd, m = _model(int(g["ngf"]), int(g["wseed"]), dev, prec)
    out = m(torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["y"]).to(dev)).cpu().numpy()
    for b in range(out.shape[0]):
        assert _rel(out[b], g["out"][b]) < TOL[prec][0], (name, prec, b, _rel(out[b], g["out"][b]))


def test_forward_strided_input_like_reference_call_site(dev):
    """test_score.py:149 feeds view_as_real(current).permute(0,3,1,2): a non-contiguous view."""
    from oracle import oracle as orc
    sd, m = _model(8, 1, dev)
    rng = np.random.default_rng(5)
    cur = torch.from_numpy(synth.cn01((7, 64, 16), rng)).to(dev) * 3.0
    xr = torch.view_as_real(cur).permute(0, 3, 1, 2)
    assert not xr.is_contiguous()
    y = torch.tensor([0, 10, 500, 1000, 1500, 2000, 2310], device=dev)
    out = m(xr, y).cpu().numpy()
    ref = orc.OracleNet(sd, 8, 64, 16).forward(xr.contiguous().cpu().numpy(), y.cpu().numpy())
    for b in range(7):
        assert _rel(out[b], ref[b]) < 2e-5


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("name", ["ald_cfg1.npz", "ald_mid.npz"])
def test_ald_matches_reference_golden_with_replayed_noise(dev, name, prec):
    """BASELINE config 1 (B=4, 2 levels x 3 steps) and a mid-trajectory case, reference noise replayed."""
    g = np.load(os.path.join(GOLDEN, name))
    sd, m = _model(int(g["ngf"]), int(g["wseed"]), dev, prec)
    _, xtol, ntol = TOL[prec]
    lv, se = g["levels"], int(g["steps_each"])
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    kw = dict(noise_var=float(g["noise_var"]), alpha_step=float(g["alpha_step"]), beta=float(g["beta"]),
              sigma_end=float(g["sigma_end"]), steps_each=se)
    X, nlog = sampler.ald_run(m, t("P"), t("Y"), t("X0"), t("H"), level_begin=int(lv[0]), level_end=int(lv[-1]) + 1,
                              ext_noise=t("ext_noise"), **kw)
    scale = np.abs(g["xs"][-1]).max()
    assert np.abs(X.cpu().numpy() - g["xs"][-1]).max() < xtol * scale
    assert np.allclose(nlog.cpu().numpy(), g["nmse"], rtol=ntol, atol=0)
    # intermediate state after the first level
    X1, _ = sampler.ald_run(m, t("P"), t("Y"), t("X0"), t("H"), level_begin=int(lv[0]), level_end=int(lv[0]) + 1,
                            ext_noise=t("ext_noise")[:se], **kw)
    assert np.abs(X1.cpu().numpy() - g["xs"][se - 1]).max() < xtol * scale


def _problem(B, Nt=64, Nr=16, Np=38, snr=0.0, seed=0):
    H = synth.cdl_like_channels(B, Nt, Nr, seed=4321 + seed)
    P = synth.qpsk_pilots(B, Nt, Np, seed=1234 + seed)
    nv = float(synth.snr_to_noise_var(snr, Nt))
    Y = synth.received_pilots(P, H, nv, seed=99 + seed)
    X0 = synth.cn01((B, Nt, Nr), np.random.default_rng(3 + seed))
    return P, Y, X0, H, nv


def test_ald_philox_noise_matches_oracle(dev):
    """Same RNG contract on both sides (include/sbc.h): no external noise needed."""
    from oracle import oracle as orc
    sd, m = _model(8, 1, dev)
    P, Y, X0, H, nv = _problem(6)
    ids = np.array([5, 0, 77, 1 << 33, 3, 9], dtype=np.uint64)
    kw = dict(noise_var=nv, alpha_step=3e-11, beta=0.01, sigma_end=SIGMA_END, level_begin=3, level_end=6,
              steps_each=2, seed=1234567891011)
    X, nlog = sampler.ald_run(m, *(torch.from_numpy(a).to(dev) for a in (P, Y, X0, H)),
                              sample_ids=torch.from_numpy(ids.astype(np.int64)).to(dev), **kw)
    Xo, nlo = orc.OracleNet(sd, 8, 64, 16).ald(P, Y, X0, H, sample_ids=ids, **kw)
    assert np.abs(X.cpu().numpy() - Xo).max() < 2e-5 * np.abs(Xo).max()
    assert np.allclose(nlog.cpu().numpy(), nlo, rtol=1e-4)


def test_per_sample_hyperparameters_and_batch_larger_than_sm_count(dev):
    """B > #SMs (CTAs loop over samples) with per-sample SNR / alpha / beta; spot-check vs the oracle;
    results must not depend on batch composition (RNG keyed by sample id)."""
    from oracle import oracle as orc
    sd, m = _model(8, 1, dev)
    B = 333
    P, Y, X0, H, _ = _problem(B)
    rng = np.random.default_rng(0)
    nv = synth.snr_to_noise_var(rng.choice([-10, 0, 10, 20, 30], B), 64).astype(np.float32)
    al = rng.choice([3e-11, 6e-11, 1e-10, 3e-10], B).astype(np.float32)
    be = rng.choice([0.1, 0.01, 0.001], B).astype(np.float32)
    kw = dict(sigma_end=SIGMA_END, level_begin=0, level_end=2, steps_each=3, seed=42)
    tt = lambda a: torch.from_numpy(a).to(dev)
    X, nlog = sampler.ald_run(m, tt(P), tt(Y), tt(X0), tt(H), noise_var=tt(nv), alpha_step=tt(al), beta=tt(be), **kw)
    sel = np.array([0, 1, 147, 148, 149, 295, 296, 332])
    Xo, nlo = orc.OracleNet(sd, 8, 64, 16).ald(P[sel], Y[sel], X0[sel], H[sel], noise_var=nv[sel], alpha_step=al[sel],
                                               beta=be[sel], sample_ids=sel.astype(np.uint64), **kw)
    Xg = X.cpu().numpy()
    assert np.abs(Xg[sel] - Xo).max() < 2e-5 * np.abs(Xo).max()
    assert np.allclose(nlog.cpu().numpy()[:, sel], nlo, rtol=1e-4)
    # the same samples in a different batch (reversed order, explicit ids) give bit-identical results
    X2, _ = sampler.ald_run(m, tt(P[sel[::-1]].copy()), tt(Y[sel[::-1]].copy()), tt(X0[sel[::-1]].copy()),
                            tt(H[sel[::-1]].copy()), noise_var=tt(nv[sel[::-1]].copy()),
                            alpha_step=tt(al[sel[::-1]].copy()), beta=tt(be[sel[::-1]].copy()),
                            sample_ids=tt(sel[::-1].astype(np.int64).copy()), **kw)
    assert np.array_equal(X2.cpu().numpy()[::-1], Xg[sel])


def test_level_range_composition_is_exact(dev):
    """[0,4) in one launch == [0,2) then [2,4) (state and RNG counters carry over exactly)."""
    sd, m = _model(8, 1, dev)
    P, Y, X0, H, nv = _problem(5, snr=10.0, seed=2)
    tt = lambda a: torch.from_numpy(a).to(dev)
    kw = dict(noise_var=nv, alpha_step=3e-11, beta=0.01, sigma_end=SIGMA_END, steps_each=3, seed=7)
    Xa, la = sampler.ald_run(m, tt(P), tt(Y), tt(X0), tt(H), level_begin=0, level_end=4, **kw)
    Xb, lb1 = sampler.ald_run(m, tt(P), tt(Y), tt(X0), tt(H), level_begin=0, level_end=2, **kw)
    Xb, lb2 = sampler.ald_run(m, tt(P), tt(Y), Xb, tt(H), level_begin=2, level_end=4, **kw)
    assert torch.equal(Xa, Xb)
    assert torch.equal(la, torch.cat([lb1, lb2]))


def test_host_buffer_abi_matches_device_abi(dev):
    sd, m = _model(8, 1, dev)
    P, Y, X0, H, nv = _problem(3, seed=4)
    tt = lambda a: torch.from_numpy(a).to(dev)
    kw = dict(noise_var=nv, alpha_step=3e-11, beta=0.01, sigma_end=SIGMA_END, level_begin=0, level_end=2,
              steps_each=3, seed=9)
    Xd, ld = sampler.ald_run(m, tt(P), tt(Y), tt(X0), tt(H), **kw)
    pm = m.packed(64, 16, dev)
    Xh = X0.copy()
    nlog = np.zeros((6, 3), np.float32)
    f = lambda v: np.full(3, v, np.float32)
    nvh, alh, beh = f(nv), f(3e-11), f(0.01)
    a = _lib.AldArgs(3, 64, 16, 38, 0, 2, 3, P.ctypes.data, Y.ctypes.data, Xh.ctypes.data, H.ctypes.data,
                     nvh.ctypes.data, alh.ctypes.data, beh.ctypes.data, SIGMA_END, nlog.ctypes.data, 9, None, None)
    _lib.check(_lib.lib().sbc_ald_run_host(pm.handle, C.byref(a)), "sbc_ald_run_host")
    assert np.array_equal(Xh, Xd.cpu().numpy()) and np.array_equal(nlog, ld.cpu().numpy())
    # forward, host variant
    x = np.random.default_rng(1).standard_normal((2, 2, 64, 16)).astype(np.float32)
    y = np.array([3, 2000], np.int64)
    out = np.empty_like(x)
    _lib.check(_lib.lib().sbc_forward_host(pm.handle, x.ctypes.data, y.ctypes.data, out.ctypes.data, 2), "fwd_host")
    ref = m(torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)).cpu().numpy()
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("prec", ["tf32x3", "tf32"])
def test_engine1_model_created_inside_the_library_is_bit_identical(dev, prec):
    """Self-contained C ABI for engine 1: sbc_model_create_from_state_ex (C++ planner, no program.py on the path) gives
    the same forward and the same ALD trajectory, bit for bit, as the model packed by the Python planner."""
    from score_based_channels_b200.engine import PackedModel
    sd, m = _model(8, 1, dev, prec)
    pm_py = m.packed(64, 16, dev)
    pm_c = PackedModel(sd, 8, 64, 16, device=dev.index or 0, precision=prec, planner="library")
    assert pm_c.prog is None and pm_c.info().engine == 1 and pm_c.info().ctas_per_sm == pm_py.info().ctas_per_sm == 2
    assert pm_c.info().n_ops == pm_py.info().n_ops and pm_c.info().arena_bytes == pm_py.info().arena_bytes
    x = np.random.default_rng(1).standard_normal((3, 2, 64, 16)).astype(np.float32)
    y = np.array([3, 2000, 2310], np.int64)
    outs = []
    for pm in (pm_py, pm_c):
        out = np.empty_like(x)
        _lib.check(_lib.lib().sbc_forward_host(pm.handle, x.ctypes.data, y.ctypes.data, out.ctypes.data, 3), "fwd_host")
        outs.append(out)
    assert np.array_equal(outs[0], outs[1])
    P, Y, X0, H, nv = _problem(3, seed=4)
    f = lambda v: np.full(3, v, np.float32)
    res = []
    for pm in (pm_py, pm_c):
        Xh, nlog = X0.copy(), np.zeros((6, 3), np.float32)
        nvh, alh, beh = f(nv), f(3e-11), f(0.01)
        a = _lib.AldArgs(3, 64, 16, 38, 0, 2, 3, P.ctypes.data, Y.ctypes.data, Xh.ctypes.data, H.ctypes.data,
                         nvh.ctypes.data, alh.ctypes.data, beh.ctypes.data, SIGMA_END, nlog.ctypes.data, 9, None, None)
        _lib.check(_lib.lib().sbc_ald_run_host(pm.handle, C.byref(a)), "sbc_ald_run_host")
        res.append((Xh, nlog))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    pm_c.close()


def test_edge_cases_and_error_codes(dev):
    sd, m = _model(8, 1, dev)
    P, Y, X0, H, nv = _problem(2)
    tt = lambda a: torch.from_numpy(a).to(dev)
    # empty batch and empty level range are no-ops
    out = m(torch.zeros((0, 2, 64, 16), device=dev), torch.zeros((0,), dtype=torch.long, device=dev))
    assert out.shape == (0, 2, 64, 16)
    X, nlog = sampler.ald_run(m, tt(P), tt(Y), tt(X0), tt(H), noise_var=nv, alpha_step=3e-11, beta=0.01,
                              sigma_end=SIGMA_END, level_begin=5, level_end=5)
    assert torch.equal(X, tt(X0)) and nlog.shape == (0, 2)
    # level range outside the schedule, Np > Nt: negative return code -> RuntimeError with a message
    with pytest.raises(RuntimeError, match="level range"):
        sampler.ald_run(m, tt(P), tt(Y), tt(X0), tt(H), noise_var=nv, alpha_step=3e-11, beta=0.01,
                        sigma_end=SIGMA_END, level_begin=0, level_end=5000)
    with pytest.raises(ValueError):
        m(torch.zeros((1, 3, 64, 16), device=dev), torch.zeros((1,), dtype=torch.long, device=dev))
    with pytest.raises(ValueError, match="multiples of 8"):
        m(torch.zeros((1, 2, 60, 16), device=dev), torch.zeros((1,), dtype=torch.long, device=dev))
    # without H there is no NMSE log but the estimate is identical
    kw = dict(noise_var=nv, alpha_step=3e-11, beta=0.01, sigma_end=SIGMA_END, level_begin=0, level_end=1, seed=1)
    Xa, la = sampler.ald_run(m, tt(P), tt(Y), tt(X0), tt(H), **kw)
    Xb, lb = sampler.ald_run(m, tt(P), tt(Y), tt(X0), None, **kw)
    assert lb is None and torch.equal(Xa, Xb)


_SUBPROCESS = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
from test_gpu_parity import _model, _problem, SIGMA_END
from score_based_channels_b200 import sampler
dev = torch.device('cuda:0')
sd, m = _model(8, 1, dev, sys.argv[2])
P, Y, X0, H, nv = _problem(3, seed=4)
tt = lambda a: torch.from_numpy(a).to(dev)
X, l = sampler.ald_run(m, tt(P), tt(Y), tt(X0), tt(H), noise_var=nv, alpha_step=3e-11, beta=0.01,
                       sigma_end=SIGMA_END, level_begin=0, level_end=2, steps_each=3, seed=9)
info = m.packed(64, 16, dev).info()
np.savez(sys.argv[1], X=X.cpu().numpy(), l=l.cpu().numpy(), smem=info.arena_in_smem, staged=info.weights_staged)
"""


@pytest.mark.parametrize("prec", ["tf32x3", "tf32"])
@pytest.mark.parametrize("env,expect", [({"SBC_STAGE_WEIGHTS": "0"}, (1, 0)), ({"SBC_FORCE_GLOBAL_ARENA": "1"}, (0, 0))])
def test_alternate_execution_modes_are_bit_identical(dev, tmp_path, env, expect, prec):
    """Parameters read straight from L2 instead of the cp.async.bulk ring, and the activation arena in
    global memory (the mode used when Nt x Nr does not fit in shared memory): same bits."""
    outs = []
    for e in ({}, env):
        f = str(tmp_path / ("o%d.npz" % len(outs)))
        subprocess.check_call([sys.executable, "-c", _SUBPROCESS % (REPO, REPO), f, prec], env={**os.environ, **e})
        outs.append(np.load(f))
    assert (int(outs[1]["smem"]), int(outs[1]["staged"])) == expect
    assert np.array_equal(outs[0]["X"], outs[1]["X"]) and np.array_equal(outs[0]["l"], outs[1]["l"])


def test_large_antenna_config_runs_from_global_arena(dev):
    """BASELINE config 5 geometry (Nt=128, Nr=32, Np=76): arena exceeds shared memory -> global arena."""
    from oracle import oracle as orc
    sd, m = _model(8, 1, dev)
    P, Y, X0, H, nv = _problem(2, Nt=128, Nr=32, Np=76)
    tt = lambda a: torch.from_numpy(a).to(dev)
    kw = dict(noise_var=nv, alpha_step=3e-11, beta=0.01, sigma_end=SIGMA_END, level_begin=0, level_end=1,
              steps_each=2, seed=5)
    X, nlog = sampler.ald_run(m, tt(P), tt(Y), tt(X0), tt(H), **kw)
    assert m.packed(128, 32, dev).info().arena_in_smem == 0
    Xo, nlo = orc.OracleNet(sd, 8, 128, 32).ald(P, Y, X0, H, **kw)
    assert np.abs(X.cpu().numpy() - Xo).max() < 2e-5 * np.abs(Xo).max()
    assert np.allclose(nlog.cpu().numpy(), nlo, rtol=1e-4)


@pytest.mark.parametrize("prec", ["tf32x3"])
def test_debug_arena_matches_schedule_simulator_op_by_op(dev, prec):
    """Localises a broken op on the GPU: arena after k ops vs the torch simulator of the same program."""
    sd, m = _model(8, 1, dev, prec)
    pm = m.packed(64, 16, dev)
    prog = pm.prog
    x = (np.random.default_rng(0).standard_normal((2, 64, 16)) * 3).astype(np.float32)
    xd = torch.from_numpy(x).to(dev)
    assert prog.park_floats > 0 and pm.info().ctas_per_sm == 2, "expected the two-CTAs-per-SM (park) plan"
    arena = torch.empty(prog.arena_floats + prog.park_floats, dtype=torch.float32, device=dev)
    for k in list(range(1, 60, 2)) + list(range(60, len(prog.ops), 5)) + list(range(len(prog.ops) - 32, len(prog.ops) + 1)):
        _lib.check(_lib.lib().sbc_debug_arena(pm.handle, xd.data_ptr(), k, arena.data_ptr(), None), "debug")
        torch.cuda.synchronize()
        _, ra = program.simulate(prog, torch.from_numpy(x), upto=k)
        op = prog.ops[k - 1]
        ga, gp = arena.cpu().numpy()[:prog.arena_floats], arena.cpu().numpy()[prog.arena_floats:]
        ra, rp = ra.numpy(), ra.park.numpy()
        if op.flags & program.F_COMPACT:
            e, r = prog.read_output(ga), prog.read_output(ra)
            assert np.abs(e - r).max() / (np.abs(r).max() + 1e-6) < 5e-5, (k - 1, op.name)
            continue
        if op.kind == program.OP_SPILL or (op.kind == program.OP_FILL and op.cin == 0):
            n = 4 * op.MT
            e, r = (gp[op.dst:op.dst + n], rp[op.dst:op.dst + n]) if op.kind == program.OP_SPILL else (ga[op.dst:op.dst + n], ra[op.dst:op.dst + n])
            assert np.abs(e - r).max() / (np.abs(r).max() + 1e-6) < 5e-5, (k - 1, op.name)
            continue
        for fi, off in enumerate((op.dst, op.acc, op.edst)):
            if off >= 0:
                eb, rb = (gp, rp) if (fi == 1 and (op.flags & program.F_ACC_G)) else (ga, ra)
                e = prog.read(eb, off, op.cout, op.oh, op.ow)
                r = prog.read(rb, off, op.cout, op.oh, op.ow)
                d, s = np.abs(e - r).max(), np.abs(r).max() + 1e-6
                assert d / s < 5e-5, (k - 1, op.name, d, s)


@pytest.mark.parametrize("Nt,Nr", [(24, 8), (40, 24)])
def test_non_power_of_two_antenna_counts_match_oracle(dev, Nt, Nr):
    """Geometries whose row widths are not powers of two (division fall-backs, direct max-pool, ragged last
    pixel tile): forward and a short ALD run against the oracle."""
    from oracle import oracle as orc
    sd, m = _model(8, 5, dev)
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((3, 2, Nt, Nr)) * 2).astype(np.float32)
    y = np.array([0, 1200, 2310])
    out = m(torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)).cpu().numpy()
    net = orc.OracleNet(sd, 8, Nt, Nr)
    ref = net.forward(x, y)
    for b in range(3):
        assert _rel(out[b], ref[b]) < 2e-5, (Nt, Nr, b, _rel(out[b], ref[b]))
    B, Np = 3, max(2, int(0.6 * Nt))
    H = synth.cdl_like_channels(B, Nt, Nr)
    P = synth.qpsk_pilots(B, Nt, Np)
    nv = float(synth.snr_to_noise_var(10.0, Nt))
    Y = synth.received_pilots(P, H, nv)
    X0 = synth.cn01((B, Nt, Nr), np.random.default_rng(4))
    kw = dict(noise_var=nv, alpha_step=3e-11, beta=0.01, sigma_end=SIGMA_END, level_begin=0, level_end=2, steps_each=3,
              seed=5)
    tt = lambda a: torch.from_numpy(a).to(dev)
    X, nlog = sampler.ald_run(m, tt(P), tt(Y), tt(X0), tt(H), **kw)
    Xo, nlo = net.ald(P, Y, X0, H, **kw)
    assert np.abs(X.cpu().numpy() - Xo).max() < 2e-5 * np.abs(Xo).max()
    assert np.allclose(nlog.cpu().numpy(), nlo, rtol=1e-4)


def test_dc_boost_and_early_stop_match_oracle(dev):
    """The approximate-MMSE knobs of reference test_mmse.py (dc_boost :25,246; target_stop :173,260-263) as
    kernel arguments: per-sample boost and per-sample last step, against the oracle; rows of the NMSE log after a
    sample's stop stay NaN and its state is the one after the stop step."""
    from oracle import oracle as orc
    sd, m = _model(8, 1, dev)
    B = 5
    P, Y, X0, H, nv = _problem(B, snr=5.0, seed=3)
    boost = np.array([1.0, 2.0, 0.5, 4.0, 1.0], np.float32)
    stop = np.array([0, 2, 5, 4, 99], np.int32)                     # 6 steps in total: 99 = no early stop
    kw = dict(noise_var=nv, alpha_step=3e-11, beta=0.01, sigma_end=SIGMA_END, level_begin=0, level_end=2, steps_each=3,
              seed=21)
    tt = lambda a: torch.from_numpy(a).to(dev)
    X, nlog = sampler.ald_run(m, tt(P), tt(Y), tt(X0), tt(H), dc_boost=tt(boost), stop_step=tt(stop), **kw)
    Xo, nlo = orc.OracleNet(sd, 8, 64, 16).ald(P, Y, X0, H, dc_boost=boost, stop_step=stop, **kw)
    assert np.abs(X.cpu().numpy() - Xo).max() < 2e-5 * np.abs(Xo).max()
    ng = nlog.cpu().numpy()
    for b in range(B):
        k = min(int(stop[b]) + 1, 6)
        assert np.allclose(ng[:k, b], nlo[:k, b], rtol=1e-4), b
        assert np.isnan(ng[k:, b]).all(), b
    # boost = 1 and no stop reproduce the plain path bit for bit
    Xa, la = sampler.ald_run(m, tt(P), tt(Y), tt(X0), tt(H), **kw)
    Xb, lb = sampler.ald_run(m, tt(P), tt(Y), tt(X0), tt(H), dc_boost=1.0, stop_step=1000, **kw)
    assert torch.equal(Xa, Xb) and torch.equal(la, lb)


@pytest.mark.parametrize("prec", ["tf32x3", "fp16x2"])
def test_ngf32_train_default_width_matches_oracle(dev, prec):
    """ngf = 32 is the default of reference train_score.py:43 (the shipped checkpoint has 8): 16x the conv work and a
    ~1.4 MB arena, so the fused kernel runs it from the L2-resident global arena; forward and one ALD step vs the
    oracle."""
    from oracle import oracle as orc
    sd, m = _model(32, 9, dev, prec)
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((2, 2, 64, 16)) * 2).astype(np.float32)
    y = np.array([10, 2000])
    out = m(torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)).cpu().numpy()
    assert m.packed(64, 16, dev).info().arena_in_smem == 0      # engine 1: L2 arena fallback; engine 2: always global
    net = orc.OracleNet(sd, 32, 64, 16)
    ref = net.forward(x, y)
    for b in range(2):
        assert _rel(out[b], ref[b]) < 2e-5, (b, _rel(out[b], ref[b]))
    P, Y, X0, H, nv = _problem(2, snr=10.0, seed=4)
    kw = dict(noise_var=nv, alpha_step=3e-11, beta=0.01, sigma_end=SIGMA_END, level_begin=0, level_end=1, steps_each=2,
              seed=3)
    tt = lambda a: torch.from_numpy(a).to(dev)
    X, nlog = sampler.ald_run(m, tt(P), tt(Y), tt(X0), tt(H), **kw)
    Xo, nlo = net.ald(P, Y, X0, H, **kw)
    assert np.abs(X.cpu().numpy() - Xo).max() < 2e-5 * np.abs(Xo).max()
    assert np.allclose(nlog.cpu().numpy(), nlo, rtol=1e-4)


@pytest.mark.parametrize("prec,tol", [("tf32x3", 2e-5), ("tf32", 5e-3)])
def test_dsm_validation_loss_matches_reference_golden_and_oracle(dev, prec, tol):
    """Fused DSM-loss launch (sbc_dsm_loss: perturb + score network + weighted L2 reduction) against the golden produced
    by the reference's own anneal_dsm_score_estimation (ncsnv2/losses/dsm.py:6-32) and against the CPU oracle."""
    from oracle import oracle as orc
    from score_based_channels_b200 import dsm
    g = np.load(os.path.join(GOLDEN, "dsm_val.npz"))
    sd, m = _model(int(g["ngf"]), int(g["wseed"]), dev, prec)
    tt = lambda a: torch.from_numpy(a).to(dev)
    per = m.dsm_losses(tt(g["samples"]), tt(g["labels"]), tt(g["z"]), float(g["anneal_power"])).cpu().numpy()
    assert np.allclose(per, g["per_sample"], rtol=tol), (per, g["per_sample"])
    ref = orc.OracleNet(sd, int(g["ngf"]), 64, 16).dsm_losses(g["samples"], g["labels"], g["z"], float(g["anneal_power"]))
    assert np.allclose(per, ref, rtol=tol)
    # the drop-in function: labels given -> the only draw is randn_like(samples), exactly as in the reference
    torch.manual_seed(5)
    loss = dsm.anneal_dsm_score_estimation(m, tt(g["samples"]), m.sigmas, tt(g["labels"]), float(g["anneal_power"]))
    torch.manual_seed(5)
    z = torch.randn_like(tt(g["samples"]))
    assert loss.dim() == 0
    assert torch.allclose(loss, m.dsm_losses(tt(g["samples"]), tt(g["labels"]), z, float(g["anneal_power"])).mean(), rtol=1e-6)
    # labels drawn inside (training-style call): finite, and the generator is consumed labels-first like the reference
    torch.manual_seed(6)
    l2 = dsm.anneal_dsm_score_estimation(m, tt(g["samples"]), m.sigmas, None, 2.)
    torch.manual_seed(6)
    lab = torch.randint(0, m.sigmas.numel(), (4,), device=dev)
    z2 = torch.randn_like(tt(g["samples"]))
    assert torch.allclose(l2, m.dsm_losses(tt(g["samples"]), lab, z2, 2.).mean(), rtol=1e-6)
    # the training call site (train mode, autograd on) is refused: there is no backward pass to hand a graph to
    m.train()
    with pytest.raises(NotImplementedError, match="forward only"):
        dsm.anneal_dsm_score_estimation(m, tt(g["samples"]), m.sigmas, None, 2.)
    with torch.no_grad():   # ... while the validation call site of train_score.py:170-185 (no_grad, any mode) works
        assert torch.isfinite(dsm.anneal_dsm_score_estimation(m, tt(g["samples"]), m.sigmas, None, 2.))
    m.eval()
    # engine 2 declines loudly instead of falling back
    m2 = make_model(sd, ngf=8, precision="fp16x2").to(dev)
    with pytest.raises(RuntimeError, match="engine-1"):
        m2.dsm_losses(tt(g["samples"]), tt(g["labels"]), tt(g["z"]), 2.)


def test_default_precision_picks_the_faster_engine_per_shape(dev):
    for ngf, want in ((8, 1), (16, 2)):
        sd = params.random_state(ngf, seed=2)
        m = make_model(sd, ngf=ngf).to(dev)          # precision=None -> "auto"
        assert m.precision == "auto"
        x = torch.randn(2, 2, 64, 16, device=dev)
        out = m(x, torch.tensor([0, 2000], device=dev))
        assert torch.isfinite(out).all() and m.packed(64, 16, dev).info().engine == want
