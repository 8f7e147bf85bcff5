"""Drop-in ``NCSNv2Deepest`` (reference ``ncsnv2/models/ncsnv2.py:198-300``).

Same constructor argument (the DotMap-style ``config``), same ``state_dict`` key names and shapes
(so ``load_state_dict(contents['model_state'])`` of a reference checkpoint works unchanged), same
``sigmas`` buffer and the same ``forward(x, y)`` contract -- but ``forward`` is one launch of the
fused sm_100a kernel through the C ABI instead of ~650 ATen kernels.  Inference only (the
reference hot path runs under ``torch.no_grad()``, ``test_score.py:150``); CUDA only.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib, params
from .engine import ALL_PRECISIONS, ENGINE2_PRECISIONS, PackedModel, engine1_runs_two_ctas_per_sm


class _Node(nn.Module):
    """Anonymous container used to reproduce the reference's nested parameter names."""


def _cfg_get(node, key, default=None):
    try:
        v = node[key] if isinstance(node, dict) else getattr(node, key)
    except (KeyError, AttributeError):
        return default
    if v is None or (hasattr(v, "__len__") and not isinstance(v, (str, bytes)) and len(v) == 0
                     and not isinstance(v, (list, tuple))):
        return default          # empty (falsy) DotMap child == "unset"
    return v


class NCSNv2Deepest(nn.Module):
    """``precision`` selects the arithmetic of the convolutions (everything else is fp32):

    * ``"fp16x2"``: engine 2 -- tcgen05 tensor cores with TMEM accumulators; every operand is an fp16 hi/lo pair
      (22 significant bits), fp32 accumulate: fp32-equivalent results;
    * ``"tf32x3"``: engine 1 -- mma.sync tensor cores, 3xTF32 error-compensated products -- fp32-equivalent
      results (the reference disables TF32, ``test_score.py:24-26``);
    * ``"tf32"``: tensor cores, plain TF32 operands (round to nearest), fp32 accumulate (fastest);
    * ``None`` / ``"auto"`` (default): the faster fp32-equivalent engine for the shape at hand, decided when the model is
      packed for an (Nt, Nr): engine 1 (``"tf32x3"``) where its two-CTAs-per-SM plan fits (ngf 8 at 64x16 and below:
      61 vs 31 est/s), engine 2 (``"fp16x2"``) elsewhere (measured on B200: 128x32 14.4 vs 13.6, ngf 16 22.5 vs 18.3,
      ngf 32 8.6 vs 3.6 est/s).
    The environment variable ``SBC_PRECISION`` overrides the default."""

    ARCH = "deepest"

    def __init__(self, config, precision: Optional[str] = None):
        super().__init__()
        self.config = config
        default = "auto" if self.ARCH == "deepest" else "fp16x2"
        self.precision = precision or os.environ.get("SBC_PRECISION", default)
        if self.precision not in ALL_PRECISIONS + ("auto",):
            raise ValueError("precision must be one of %s or 'auto'" % (ALL_PRECISIONS,))
        if self.ARCH != "deepest" and self.precision not in ENGINE2_PRECISIONS:
            raise NotImplementedError("%s runs on engine 2 only (precision 'fp16x2'); the engine-1 planner knows "
                                      "NCSNv2Deepest" % type(self).__name__)
        model, data = config.model, config.data
        self.ngf = int(model.ngf)
        self.num_classes = int(model.num_classes)
        self.channels = int(_cfg_get(data, "channels", 2))
        norm = _cfg_get(model, "normalization", "InstanceNorm++")
        act = str(_cfg_get(model, "nonlinearity", "elu")).lower()
        if norm != "InstanceNorm++" or act != "elu":
            raise NotImplementedError("fused kernel implements InstanceNorm++ / ELU (the shipped configuration, "
                                      "train_score.py:39-40); got %r / %r" % (norm, act))
        # reference ncsnv2.py:201-202,270-271: both unset -> h = 2x - 1 is applied
        if _cfg_get(data, "logit_transform", False) or _cfg_get(data, "rescaled", False):
            raise NotImplementedError("logit_transform / rescaled inputs are not used by score_based_channels")
        if str(_cfg_get(model, "sigma_dist", "geometric")) != "geometric":
            raise NotImplementedError("only the geometric sigma schedule is used on this path")
        sig = params.get_sigmas_np(float(model.sigma_begin), float(model.sigma_end), self.num_classes)
        self.register_buffer("sigmas", torch.from_numpy(sig))
        # parameters under the reference's names
        rng = np.random.default_rng(0)
        init = params.random_state(self.ngf, self.channels, self.num_classes, float(model.sigma_begin),
                                   float(model.sigma_end), seed=int(rng.integers(1 << 30)), arch=self.ARCH)
        for name, shape in params.param_shapes(self.ngf, self.channels, self.num_classes, self.ARCH).items():
            if name == "sigmas":
                continue
            node = self
            parts = name.split(".")
            for p in parts[:-1]:
                if p not in node._modules:
                    node.add_module(p, _Node())
                node = node._modules[p]
            node.register_parameter(parts[-1], nn.Parameter(torch.from_numpy(init[name]), requires_grad=False))
        self._packed: Optional[PackedModel] = None
        self._packed_key = None

    # ------------------------------------------------------------------
    def _key(self, Nt, Nr, device):
        return (Nt, Nr, device.index if device.index is not None else torch.cuda.current_device(), self.precision,
                tuple(p._version for p in self.parameters()), self.sigmas._version,
                tuple(p.data_ptr() for p in self.parameters()))

    def packed(self, Nt: int, Nr: int, device: torch.device) -> PackedModel:
        """(Re)pack the current parameters for an Nt x Nr input on `device` (cached)."""
        key = self._key(Nt, Nr, device)
        if self._packed is None or self._packed_key != key:
            if self._packed is not None:
                self._packed.close()
                self._packed, self._packed_key = None, None
            state = {k: v.detach().cpu().numpy() for k, v in self.state_dict().items()}
            prec = self.precision
            if prec == "auto":
                prec = "tf32x3" if engine1_runs_two_ctas_per_sm(state, self.ngf, Nt, Nr, self.channels) else "fp16x2"
            self._packed = PackedModel(state, self.ngf, Nt, Nr, key[2], self.channels, prec)
            self._packed_key = key
        return self._packed

    def forward(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """score = f(2x-1)/sigmas[y]; x fp32 [B,channels,Nt,Nr] (any strides), y int64 [B]."""
        if not x.is_cuda:
            raise RuntimeError("NCSNv2Deepest (B200) runs on CUDA tensors only; there is no CPU path")
        if x.dim() != 4 or x.shape[1] != self.channels:
            raise ValueError("expected x of shape [B,%d,Nt,Nr], got %s" % (self.channels, tuple(x.shape)))
        if x.dtype != torch.float32:
            x = x.float()
        B, _, Nt, Nr = x.shape
        y = y.to(device=x.device, dtype=torch.int64).contiguous()
        if y.numel() != B:
            raise ValueError("labels must have one entry per sample")
        pm = self.packed(Nt, Nr, x.device)
        out = torch.empty((B, self.channels, Nt, Nr), dtype=torch.float32, device=x.device)
        if B == 0:
            return out
        strides = (C.c_int64 * 4)(*x.stride())
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().sbc_forward(pm.handle, x.data_ptr(), strides, y.data_ptr(), out.data_ptr(), B,
                                              C.c_void_p(stream)), "sbc_forward")
        return out


    def dsm_losses(self, samples: torch.Tensor, labels: torch.Tensor, z: torch.Tensor, anneal_power: float = 2.) -> torch.Tensor:
        """Per-sample denoising-score-matching losses [B] for the clean ``samples`` perturbed by ``z * sigmas[labels]``
        (reference ``ncsnv2/losses/dsm.py:13-31`` without the final mean), one fused launch, no autograd."""
        if not samples.is_cuda:
            raise RuntimeError("NCSNv2Deepest (B200) runs on CUDA tensors only; there is no CPU path")
        if samples.dim() != 4 or samples.shape[1] != self.channels or z.shape != samples.shape:
            raise ValueError("expected samples and z of shape [B,%d,Nt,Nr]" % self.channels)
        B, _, Nt, Nr = samples.shape
        x = samples.float().contiguous()
        zz = z.to(device=x.device, dtype=torch.float32).contiguous()
        y = labels.to(device=x.device, dtype=torch.int64).contiguous()
        if y.numel() != B:
            raise ValueError("labels must have one entry per sample")
        pm = self.packed(Nt, Nr, x.device)
        out = torch.empty((B,), dtype=torch.float32, device=x.device)
        if B == 0:
            return out
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().sbc_dsm_loss(pm.handle, x.data_ptr(), y.data_ptr(), zz.data_ptr(), float(anneal_power),
                                               out.data_ptr(), B, C.c_void_p(stream)), "sbc_dsm_loss")
        return out


class NCSNv2Deeper(NCSNv2Deepest):
    """Drop-in ``NCSNv2Deeper`` (reference ``ncsnv2/models/ncsnv2.py:94-195``): five encoder stages, two mean-pools, dilated
    stages at a quarter of the input resolution.  Same forward contract; runs on engine 2 (tcgen05), whose C++ planner reads
    the architecture off the state-dict keys."""
    ARCH = "deeper"


class NCSNv2(NCSNv2Deepest):
    """Drop-in ``NCSNv2`` (reference ``ncsnv2/models/ncsnv2.py:11-91``): four encoder stages, one mean-pool, dilated stages
    at half the input resolution.  Same forward contract; runs on engine 2 (tcgen05)."""
    ARCH = "ncsnv2"
