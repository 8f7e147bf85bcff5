"""Drop-in for ``ncsnv2.losses.dsm.anneal_dsm_score_estimation`` (reference ``ncsnv2/losses/dsm.py:6-32``), forward only.

The reference evaluates this function in two places of ``train_score.py``: with autograd for the training step
(``:145-153``) and under ``torch.no_grad()`` for the validation loss (``:170-185``).  This module covers the second use:
the perturbation, the score network and the weighted L2 reduction run in ONE fused launch of the engine-1 kernel
(``sbc_dsm_loss``).  The random draws are made with torch in the reference's order (labels first, then ``randn_like``), so
a seeded run consumes the generator exactly like the reference does.  There is no backward pass."""
from __future__ import annotations

import torch


def anneal_dsm_score_estimation(scorenet, samples, sigmas, labels=None, anneal_power=2.):
    if getattr(scorenet, "training", False) and torch.is_grad_enabled():   # the training call site (train_score.py:145-153)
        raise NotImplementedError("the B200 library evaluates the DSM loss forward only (validation loss); "
                                  "the training backward pass is not implemented")
    if len(sigmas) != scorenet.sigmas.numel():
        raise ValueError("sigmas must be the schedule the score network was built with")
    if labels is None:   # dsm.py:9-12
        labels = torch.randint(0, len(sigmas), (samples.shape[0],), device=samples.device)
    z = torch.randn_like(samples)   # dsm.py:15: noise = randn_like(samples) * used_sigmas
    return scorenet.dsm_losses(samples, labels, z, anneal_power).mean(dim=0)
