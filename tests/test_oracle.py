"""Pins the CPU oracle (oracle/sbc_oracle.c) against golden vectors produced by the reference's own
modules (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from score_based_channels_b200 import params

from conftest import GOLDEN


def _rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))


@pytest.mark.parametrize("name,H,W", [("forward_ngf8.npz", 64, 16), ("forward_ngf8_32x8.npz", 32, 8),
                                      ("forward_ngf16.npz", 64, 16)])
def test_forward_matches_reference_golden(name, H, W):
    g = np.load(os.path.join(GOLDEN, name))
    sd = params.random_state(int(g["ngf"]), seed=int(g["wseed"]))
    net = orc.OracleNet(sd, int(g["ngf"]), H, W)
    out = net.forward(g["x"], g["y"])
    for b in range(out.shape[0]):
        assert _rel(out[b], g["out"][b]) < 2e-5, (name, b)


@pytest.mark.parametrize("name", ["ald_cfg1.npz", "ald_mid.npz"])
def test_ald_matches_reference_golden(name):
    """BASELINE config 1 (B=4, 2 levels x 3 steps) and a mid-trajectory case, replaying the reference's noise."""
    g = np.load(os.path.join(GOLDEN, name))
    sd = params.random_state(int(g["ngf"]), seed=int(g["wseed"]))
    net = orc.OracleNet(sd, int(g["ngf"]), 64, 16)
    lv = g["levels"]
    steps_each = int(g["steps_each"])
    X, nlog = net.ald(g["P"], g["Y"], g["X0"], g["H"], noise_var=float(g["noise_var"]),
                      alpha_step=float(g["alpha_step"]), beta=float(g["beta"]), sigma_end=float(g["sigma_end"]),
                      level_begin=int(lv[0]), level_end=int(lv[-1]) + 1, steps_each=steps_each,
                      ext_noise=g["ext_noise"])
    scale = np.abs(g["xs"][-1]).max()
    assert np.abs(X - g["xs"][-1]).max() < 1e-5 * scale
    assert np.allclose(nlog, g["nmse"], rtol=2e-5, atol=0)
    # prefix property: running only the first level reproduces the intermediate state
    X1, _ = net.ald(g["P"], g["Y"], g["X0"], g["H"], noise_var=float(g["noise_var"]),
                    alpha_step=float(g["alpha_step"]), beta=float(g["beta"]), sigma_end=float(g["sigma_end"]),
                    level_begin=int(lv[0]), level_end=int(lv[0]) + 1, steps_each=steps_each,
                    ext_noise=g["ext_noise"][:steps_each])
    assert np.abs(X1 - g["xs"][steps_each - 1]).max() < 1e-5 * scale


def test_philox_noise_statistics_and_determinism():
    a = orc.noise(seed=5, sid=3, step=11, n_elem=1024)
    b = orc.noise(seed=5, sid=3, step=11, n_elem=1024)
    assert np.array_equal(a, b)
    assert not np.array_equal(a, orc.noise(5, 4, 11, 1024))
    assert not np.array_equal(a, orc.noise(5, 3, 12, 1024))
    z = np.concatenate([orc.noise(1, s, 0, 1024) for s in range(64)])
    assert abs(np.mean(np.abs(z) ** 2) - 1.0) < 0.02          # unit power: re, im ~ N(0, 1/2)
    assert abs(np.var(z.real) - 0.5) < 0.02 and abs(np.var(z.imag) - 0.5) < 0.02
    assert abs(np.mean(z.real)) < 0.02 and abs(np.mean(z.real * z.imag)) < 0.02
    # odd element counts are handled
    assert np.array_equal(orc.noise(5, 3, 11, 1023), a[:1023])


LOCAL_CKPT = os.path.join(os.path.dirname(GOLDEN), os.pardir, "fixtures_local", "score-deepest-cdl-c.pt")


@pytest.mark.skipif(not os.path.exists(LOCAL_CKPT), reason="fixtures_local/ (reference checkpoint) not present")
def test_oracle_matches_reference_on_shipped_checkpoint():
    """Shipped weights: the oracle's error against the reference's fp64 output stays within 3x the
    reference's own fp32 error (the output is ill-conditioned at small sigma, see make_golden.py)."""
    from score_based_channels_b200 import entry_common as ec
    g = np.load(os.path.join(GOLDEN, "real_ckpt_forward.npz"))
    sd = {k: v.numpy() for k, v in ec.load_checkpoint(LOCAL_CKPT)["model_state"].items()}
    out = orc.OracleNet(sd, 8, 64, 16).forward(g["x"], g["y"]).astype(np.float64)
    for b in range(out.shape[0]):
        e32, e = _rel(g["out32"][b].astype(np.float64), g["out64"][b]), _rel(out[b], g["out64"][b])
        assert e <= 3 * e32 + 2e-6, (b, e, e32)


def test_oracle_dsm_loss_matches_reference_function():
    """anneal_dsm_score_estimation (reference ncsnv2/losses/dsm.py:6-32, evaluated by the reference itself under no_grad
    as train_score.py:170-185 does) vs its CPU restatement: per-sample terms and the mean."""
    g = np.load(os.path.join(GOLDEN, "dsm_val.npz"))
    sd = params.random_state(int(g["ngf"]), seed=int(g["wseed"]))
    net = orc.OracleNet(sd, int(g["ngf"]), 64, 16)
    per = net.dsm_losses(g["samples"], g["labels"], g["z"], float(g["anneal_power"]))
    assert np.allclose(per, g["per_sample"], rtol=5e-6)
    assert abs(float(per.mean()) - float(g["loss"])) < 5e-6 * float(g["loss"])
