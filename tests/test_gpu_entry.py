"""Entry-point drop-ins on the GPU (need the reference checkpoint + sample data, which travel to the GPU box
in the git-ignored fixtures_local/; skipped when absent)."""
import os

import numpy as np
import pytest
import torch

from conftest import REPO

pytestmark = pytest.mark.gpu

FX = os.path.join(REPO, "fixtures_local")
CKPT = os.path.join(FX, "score-deepest-cdl-c.pt")
MAT = os.path.join(FX, "CDL-C_Nt64_Nr16_ULA0.50_seed4321.mat")
need = pytest.mark.skipif(not (os.path.exists(CKPT) and os.path.exists(MAT)), reason="fixtures_local/ not present")


@need
def test_test_score_entry_point_small(tmp_path, monkeypatch):
    monkeypatch.chdir(REPO)
    from score_based_channels_b200 import test_score
    out = test_score.main(["--ckpt", CKPT, "--out_dir", str(tmp_path), "--levels", "3", "--num_channels", "6",
                           "--seed", "0", "--no_plot"])
    res = torch.load(os.path.join(str(tmp_path), "results.pt"), weights_only=False)
    assert set(res.keys()) == {"nmse_log", "avg_nmse", "best_nmse", "spacing_range", "pilot_alpha_range", "snr_range",
                               "val_config"}                                   # test_score.py:192-199
    assert res["nmse_log"].shape == (1, 1, 17, 9, 6) and res["nmse_log"].dtype == np.float64
    assert res["avg_nmse"].shape == (1, 1, 17, 9) and res["best_nmse"].shape == (1, 1, 17)
    assert np.all(np.isfinite(res["nmse_log"])) and np.all(res["nmse_log"] > 0)
    assert np.allclose(res["snr_range"], np.arange(-10, 32.5, 2.5))
    assert res["val_config"].data.num_pilots == 38
    # a random CN(0,1) start against a unit-power channel: NMSE ~ 2 at the first step
    assert 1.0 < res["avg_nmse"][0, 0, :, 0].mean() < 4.0


@need
def test_tune_hparams_entry_point_small(tmp_path, monkeypatch):
    monkeypatch.chdir(REPO)
    from score_based_channels_b200 import tune_hparams_score
    tune_hparams_score.main(["--ckpt", CKPT, "--out_dir", str(tmp_path), "--levels", "2", "--num_channels", "4",
                             "--alpha_step_range", "3e-11", "1e-10", "--beta_noise_range", "0.1", "0.01", "0.001",
                             "--seed", "0", "--no_plot"])
    res = torch.load(os.path.join(str(tmp_path), "CDL-C-hyperparameters.pt"), weights_only=False)
    assert set(res.keys()) == {"nmse_log", "avg_nmse", "best_nmse", "best_alpha_snr", "best_beta_snr", "snr_range",
                               "alpha_step_range", "beta_noise_range", "config", "args"}   # tune_hparams_score.py:180-188
    assert res["nmse_log"].shape == (2, 3, 17, 6, 4)
    assert len(res["best_alpha_snr"]) == 17 and all(a in (3e-11, 1e-10) for a in res["best_alpha_snr"])
    assert np.all(np.isfinite(res["nmse_log"]))


@need
def test_real_checkpoint_forward_matches_reference_golden():
    """Shipped weights (ngf=8), shipped channels, six noise levels: fused forward vs the reference module's own
    output (tests/golden/real_ckpt_forward.npz, made by tests/golden/make_golden.py).  With trained weights the
    output at small sigma is an ill-conditioned difference: the reference's OWN fp32 output deviates from its
    fp64 output by e32 = 1.4e-6 ... 3.5e-4 (relative L2).  Tolerances, per sample, on the error vs the fp64
    reference: fp32-equivalent mode (tf32x3) <= 3 * e32 + 2e-6; TF32 mode <= 4096 * e32 (unit roundoff ratio
    2^13, half of it) and <= 6e-3 wherever the problem is well conditioned (e32 < 5e-6)."""
    from conftest import GOLDEN
    from score_based_channels_b200 import entry_common as ec
    dev = torch.device("cuda:0")
    g = np.load(os.path.join(GOLDEN, "real_ckpt_forward.npz"))
    contents = ec.load_checkpoint(CKPT)
    rel = lambda a, b: float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))
    for prec in ("tf32x3", "fp16x2", "tf32"):
        m = ec.build_model(contents["config"], contents["model_state"], dev, prec)
        out = m(torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["y"]).to(dev)).cpu().numpy().astype(np.float64)
        for b in range(out.shape[0]):
            e32, e = rel(g["out32"][b].astype(np.float64), g["out64"][b]), rel(out[b], g["out64"][b])
            if prec in ("tf32x3", "fp16x2"):
                assert e <= 3 * e32 + 2e-6, (prec, b, e, e32)
            else:
                assert e <= 4096 * e32, (prec, b, e, e32)
                if e32 < 5e-6:
                    assert e < 6e-3, (prec, b, e)


@need
def test_tf32_mode_trajectory_nmse_within_1e3_of_fp32_equivalent_mode():
    """north_star tolerance for the fast mode: with the SHIPPED checkpoint, shipped channels and identical
    pilots / noise (same Philox seed), the per-channel NMSE of the TF32 mode stays within 1e-3 (linear units)
    of the fp32-equivalent mode at every logged step of a sub-sampled full-range schedule and at the end --
    although the TF32 *forward* is off by tens of percent at the smallest sigmas (the update is alpha-weighted
    and the data term makes the dynamics contractive, SURVEY.md section 7 hard part 3)."""
    from score_based_channels_b200 import entry_common as ec, hdf5_min, sampler, synth
    dev = torch.device("cuda:0")
    contents = ec.load_checkpoint(CKPT)
    h = hdf5_min.loadmat_v73(MAT)["output_h"][:32, 0].astype(np.complex64)
    Hn = (np.conj(np.transpose(h, (0, 2, 1))) / 0.363263).astype(np.complex64)
    B, Nt, Nr, Np = Hn.shape[0], 64, 16, 38
    P = synth.qpsk_pilots(B, Nt, Np, seed=1)
    snr = np.array([0.0, 10.0, 20.0, 30.0])[np.arange(B) % 4]
    nv = synth.snr_to_noise_var(snr, Nt).astype(np.float32)
    Y = synth.received_pilots(P, Hn, nv, seed=2)
    X0 = synth.cn01((B, Nt, Nr), np.random.default_rng(3))
    tt = lambda a: torch.from_numpy(a).to(dev)
    res = {}
    for prec in ("tf32x3", "tf32"):
        m = ec.build_model(contents["config"], contents["model_state"], dev, prec)
        X = tt(X0)
        logs = []
        # every 7th level across the whole schedule (331 levels x 3 steps), chained launches, same seed
        for lvl in range(0, 2311, 7):
            X, nl = sampler.ald_run(m, tt(P), tt(Y), X, tt(Hn), noise_var=tt(nv), alpha_step=3e-11, beta=0.01,
                                    level_begin=lvl, level_end=lvl + 1, steps_each=3, seed=77)
            logs.append(nl)
        res[prec] = torch.cat(logs).cpu().numpy()
    d = np.abs(res["tf32"] - res["tf32x3"])
    assert np.isfinite(res["tf32"]).all() and np.isfinite(res["tf32x3"]).all()
    assert d[-1].max() < 1e-3, ("final", d[-1].max())
    assert d.max() < 1e-3 * max(1.0, res["tf32x3"].max()), ("trajectory", d.max())
    # sanity: the sampler is estimating (even on this 7x sub-sampled schedule NMSE at 30 dB is far below 0 dB)
    assert res["tf32x3"][-1, snr == 30.0].mean() < 0.5 * res["tf32x3"][-1, snr == 0.0].mean()


@need
def test_test_mmse_entry_point_small(tmp_path, monkeypatch):
    """Approximate-MMSE drop-in (reference test_mmse.py): result-file keys / shapes, early stop, start points."""
    monkeypatch.chdir(REPO)
    from score_based_channels_b200 import test_mmse
    hyper = os.path.join(str(tmp_path), "hyper.pt")
    torch.save({"best_step_idx": np.full((1, 2), 3e-11), "best_noise_idx": np.full((1, 2), 0.01),
                "best_stop_idx": np.array([[4, 100]])}, hyper)
    for sp in ("Noise", "Adjoint", "LS"):
        out = test_mmse.main(["--gpu", "0", "--ckpt", CKPT, "--hyper", hyper, "--out_dir", str(tmp_path), "--levels", "3",
                              "--kept_samples", "4", "--mmse_avg", "3", "--snr_range", "0", "10", "--start_point", sp,
                              "--dc_boost", "1.5", "--seed", "0"])
        res = torch.load(os.path.join(str(tmp_path), "model_CDL-C_channel_CDL-C.pt"), weights_only=False)
        assert set(res.keys()) == {"spacing_range", "pilot_alpha_range", "args", "config", "snr_range", "val_config",
                                   "oracle_log", "oracle_H", "saved_H", "mmse_nmse"}          # test_mmse.py:278-288 (+1)
        assert res["oracle_log"].shape == (1, 1, 2, 9, 4, 3) and res["saved_H"].shape == (1, 1, 2, 4, 3, 64, 16)
        lg = res["oracle_log"][0, 0]
        assert np.isfinite(lg[0, :5]).all() and np.isnan(lg[0, 5:]).all()     # SNR 0: stopped after step 4
        assert np.isfinite(lg[1]).all()                                       # SNR 1: stop beyond the schedule
        assert np.isfinite(res["mmse_nmse"]).all() and (res["mmse_nmse"] > 0).all()
        assert res["oracle_H"].shape == (4, 64, 16)
    # averaging posterior samples cannot be worse than the mean of the individual NMSEs (Jensen), SNR 1, last step
    assert (out["mmse_nmse"][0, 0, 1] <= np.nanmean(out["oracle_log"][0, 0, 1, -1], axis=-1) * (1 + 1e-5)).all()
