"""Convenience constructors around the drop-in ``NCSNv2Deepest``."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from . import dotmap_shim
from .ncsnv2 import NCSNv2Deepest


def make_config(ngf: int = 8, num_classes: int = 2311, sigma_begin: float = 27.77,
                sigma_end: float = 2.599515446446343e-4, Nt: int = 64, Nr: int = 16):
    """A config tree with the fields the reference reads (train_score.py:34-67; Appendix A of SURVEY.md)."""
    DotMap = dotmap_shim.DotMap
    cfg = DotMap()
    cfg.device = "cuda:0"
    cfg.model.ngf = ngf
    cfg.model.num_classes = num_classes
    cfg.model.normalization = "InstanceNorm++"
    cfg.model.nonlinearity = "elu"
    cfg.model.sigma_dist = "geometric"
    cfg.model.sigma_begin = sigma_begin
    cfg.model.sigma_end = sigma_end
    cfg.data.channels = 2
    cfg.data.image_size = [Nr, Nt]
    cfg.sampling.steps_each = 3
    return cfg


def make_model(state: Dict[str, np.ndarray], ngf: int = 8, precision=None, **cfg_kw) -> NCSNv2Deepest:
    """NCSNv2Deepest carrying `state` (reference-keyed numpy arrays; sigmas define the schedule)."""
    sig = np.asarray(state["sigmas"], dtype=np.float64)
    cfg = make_config(ngf=ngf, num_classes=sig.size, **cfg_kw)
    m = NCSNv2Deepest(cfg, precision=precision)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v, dtype=np.float32)) for k, v in state.items()})
    return m.eval()
