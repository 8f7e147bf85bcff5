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
def test_real_checkpoint_forward_matches_oracle():
    """Shipped weights (ngf=8), shipped channels: fused forward vs the CPU oracle at three noise levels."""
    from oracle import oracle as orc
    from score_based_channels_b200 import entry_common as ec, hdf5_min
    dev = torch.device("cuda:0")
    contents = ec.load_checkpoint(CKPT)
    sd = {k: v.numpy() for k, v in contents["model_state"].items()}
    h = hdf5_min.loadmat_v73(MAT)["output_h"][:6, 0].astype(np.complex64)
    Hn = np.conj(np.transpose(h, (0, 2, 1))) / 0.363263
    labels = np.array([0, 500, 1000, 1500, 2000, 2310])
    rng = np.random.default_rng(0)
    x = Hn + sd["sigmas"][labels][:, None, None] * ((rng.standard_normal(Hn.shape) + 1j * rng.standard_normal(Hn.shape)) / np.sqrt(2))
    xr = np.stack([x.real, x.imag], 1).astype(np.float32)
    ref = orc.OracleNet(sd, 8, 64, 16).forward(xr, labels)
    for prec, tol in (("tf32x3", 2e-5), ("tf32", 6e-3)):
        m = ec.build_model(contents["config"], contents["model_state"], dev, prec)
        out = m(torch.from_numpy(xr).to(dev), torch.from_numpy(labels).to(dev)).cpu().numpy()
        for b in range(6):
            rel = np.linalg.norm(out[b] - ref[b]) / np.linalg.norm(ref[b])
            assert rel < tol, (prec, b, rel)
