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


class Channels(Dataset):
    """MIMO Channels"""

    def __init__(self, seed, config, norm=None):
        target_spacings = config.data.spacing_list
        target_channel = config.data.channel
        self.channels = []
        self.spacings = np.copy(target_spacings)
        self.filenames = []
        for spacing in target_spacings:
            filename = find_data_file(target_channel, spacing, seed)
            self.filenames.append(filename)
            contents = loadmat_v73(filename)
            channels = np.asarray(contents["output_h"], dtype=np.complex64)
            self.channels.append(channels[:, 0])            # first subcarrier of each symbol (loaders.py:33)
        self.channels = np.asarray(self.channels)
        self.channels = np.reshape(self.channels, (-1, self.channels.shape[-2], self.channels.shape[-1]))

        if type(norm) == list:
            self.mean, self.std = norm[0], norm[1]
        elif norm == "entrywise":
            self.mean = np.mean(self.channels, axis=0)
            self.std = np.std(self.channels, axis=0)
        elif norm == "global":
            self.mean = 0.
            self.std = np.std(self.channels)

        # random QPSK pilots (loaders.py:52-55)
        shape = (self.channels.shape[0], config.data.image_size[1], config.data.num_pilots)
        self.pilots = 1 / np.sqrt(2) * (2 * np.random.binomial(1, 0.5, size=shape) - 1 +
                                        1j * (2 * np.random.binomial(1, 0.5, size=shape) - 1))
        self.noise_power = 1 / np.sqrt(2) * config.data.noise_std

    def __len__(self):
        return len(self.channels)

    def __getitem__(self, idx):
        if torch.is_tensor(idx):
            idx = idx.tolist()
        H_cplx = self.channels[idx]
        H_cplx_norm = (H_cplx - self.mean) / self.std
        H_real_norm = np.stack((np.real(H_cplx_norm), np.imag(H_cplx_norm)), axis=0)
        P = self.pilots[idx]
        Y = np.matmul(H_cplx, P)
        N = self.noise_power * (np.random.normal(size=Y.shape) + 1j * np.random.normal(size=Y.shape))
        Y = Y + N
        eigvals = np.real(np.linalg.eigvals(np.matmul(P, np.conj(P.T))))
        H_herm = np.conj(np.transpose(H_cplx))
        H_herm_norm = np.conj(np.transpose(H_cplx_norm))
        H_real_herm_norm = np.stack((np.real(H_herm_norm), np.imag(H_herm_norm)), axis=0)
        P_herm = np.conj(np.transpose(P))
        Y_herm = np.conj(np.transpose(Y))
        return {"H": H_real_norm.astype(np.float32),
                "H_herm": H_real_herm_norm.astype(np.float32),
                "H_herm_cplx": H_herm.astype(np.complex64),
                "P": self.pilots[idx].astype(np.complex64),
                "P_herm": P_herm.astype(np.complex64),
                "Y": Y.astype(np.complex64),
                "Y_herm": Y_herm.astype(np.complex64),
                "eig1": eigvals[0].astype(np.float32),
                "sigma_n": np.float32(self.noise_power),
                "idx": int(idx)}
