"""``Channels`` dataset -- drop-in for reference ``src/score_based_channels/loaders.py:8-107``.

Same constructor ``Channels(seed, config, norm)``, same attributes (``channels, mean, std, pilots,
noise_power, filenames``) and the same per-item dict.  The MATLAB v7.3 file is read with the bundled
minimal HDF5 reader (``hdf5storage`` is not available here).  Like the reference, pilots and the
per-item noise are drawn from numpy's *global* RNG (unseeded unless the caller seeds it).

File lookup: ``./data/<channel>_Nt64_Nr16_ULA<spacing>_seed<seed>.mat`` exactly as the reference
(``loaders.py:23-24``), then ``$SBC_DATA_DIR``, ``./fixtures_local`` and ``./sample_data``.  The training-set
file (seed 1234) that ``test_score.py:66-69`` uses only for its ``.std`` is not shipped with the
reference; when it is absent the validation file is used for the normalisation statistics (a warning is
printed; SURVEY.md section 8, deviation 10).
"""
from __future__ import annotations

import os
import warnings

import numpy as np
import torch
from torch.utils.data import Dataset

from .hdf5_min import loadmat_v73

_SEARCH = ("./data", os.environ.get("SBC_DATA_DIR", ""), "./fixtures_local", "./sample_data",
           os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fixtures_local"))


def find_data_file(channel: str, spacing: float, seed: int, allow_other_seed: bool = True) -> str:
    names = ["%s_Nt64_Nr16_ULA%.2f_seed%d.mat" % (channel, spacing, seed)]
    for d in _SEARCH:
        if d and os.path.exists(os.path.join(d, names[0])):
            return os.path.join(d, names[0])
    if allow_other_seed:
        for d in _SEARCH:
            if not d or not os.path.isdir(d):
                continue
            for f in sorted(os.listdir(d)):
                if f.startswith("%s_Nt64_Nr16_ULA%.2f_seed" % (channel, spacing)) and f.endswith(".mat"):
                    warnings.warn("data file %s not found; using %s for the normalisation statistics"
                                  % (names[0], os.path.join(d, f)))
                    return os.path.join(d, f)
    raise FileNotFoundError("./data/%s (also searched $SBC_DATA_DIR, ./fixtures_local, ./sample_data)" % names[0])


def _herm(a: np.ndarray) -> np.ndarray:
    return np.conj(a.T)


def _planes(a: np.ndarray) -> np.ndarray:
    """complex [r, c] -> float32 [2, r, c] (re, im): the network's channel layout."""
    return np.stack((a.real, a.imag), axis=0).astype(np.float32)


class Channels(Dataset):
    """Validation / training channel set with the interface of the reference ``Channels`` (``loaders.py:8-107``).

    Behavioural contract kept from the reference (restated, not copied):
    * one ``.mat`` per antenna spacing in ``config.data.spacing_list``; of every realisation only the first
      subcarrier/symbol slice ``output_h[:, 0]`` ([Nr, Nt]) is used (``loaders.py:30-33``);
    * ``norm``: ``[mean, std]`` list = use as given; ``'global'`` = mean 0, std of all entries; ``'entrywise'`` =
      per-entry statistics (``loaders.py:39-49``);
    * QPSK pilots ``(+-1 +-1j)/sqrt(2)`` of shape [N, Nt, num_pilots], real parts drawn before imaginary parts from
      numpy's global RNG (``loaders.py:52-55``), per-item measurement noise likewise (``:79-82``)."""

    _ITEM_KEYS = ("H", "H_herm", "H_herm_cplx", "P", "P_herm", "Y", "Y_herm", "eig1", "sigma_n", "idx")

    def __init__(self, seed, config, norm=None, allow_other_seed=True):
        data = config.data
        self.spacings = np.array(data.spacing_list, copy=True)
        self.filenames = [find_data_file(data.channel, sp, seed, allow_other_seed) for sp in data.spacing_list]
        per_file = [np.asarray(loadmat_v73(f)["output_h"], dtype=np.complex64)[:, 0] for f in self.filenames]
        stacked = np.asarray(per_file)
        self.channels = stacked.reshape((-1,) + stacked.shape[-2:])

        if isinstance(norm, list):
            self.mean, self.std = norm
        elif norm == "global":
            self.mean, self.std = 0., np.std(self.channels)
        elif norm == "entrywise":
            self.mean, self.std = np.mean(self.channels, axis=0), np.std(self.channels, axis=0)

        n_tx = data.image_size[1]
        bits = [np.random.binomial(1, 0.5, size=(len(self.channels), n_tx, data.num_pilots)) for _ in range(2)]
        self.pilots = ((2 * bits[0] - 1) + 1j * (2 * bits[1] - 1)) / np.sqrt(2)
        self.noise_power = data.noise_std / np.sqrt(2)

    def __len__(self):
        return self.channels.shape[0]

    def __getitem__(self, idx):
        idx = idx.tolist() if torch.is_tensor(idx) else idx
        h = self.channels[idx]                              # [Nr, Nt]
        hn = (h - self.mean) / self.std
        p = self.pilots[idx]                                # [Nt, Np]
        y = h @ p
        noise = [np.random.normal(size=y.shape) for _ in range(2)]
        y = y + self.noise_power * (noise[0] + 1j * noise[1])
        gram_eig = np.linalg.eigvals(p @ _herm(p)).real
        values = (_planes(hn), _planes(_herm(hn)), _herm(h).astype(np.complex64), p.astype(np.complex64),
                  _herm(p).astype(np.complex64), y.astype(np.complex64), _herm(y).astype(np.complex64),
                  gram_eig[0].astype(np.float32), np.float32(self.noise_power), int(idx))
        return dict(zip(self._ITEM_KEYS, values))
