"""On-device evaluation of a training ensemble: the quantities the reference computes per exported
dictionary on the CPU — FVU (standard_metrics.py:310-314), mean L0 (:305-308), features ever active (:446-454) —
obtained for all M models at once from the engine's forward pass (no parameter update, no dense fp32 code unless
``n_ever_active`` is requested).

FVU is taken in the space the model reconstructs (the centred space for FunctionalTiedSAE with a non-trivial
centring: an orthogonal rotation leaves it unchanged, a non-uniform ``center_scale`` does not)."""
from __future__ import annotations

from typing import Dict

import torch


def evaluate(ensemble, batch: torch.Tensor, n_ever_active: bool = False) -> Dict[str, torch.Tensor]:
    """batch [B, d] (CUDA or pinned host). Returns per-model tensors on the ensemble's device:
    ``fvu``, ``mean_l0``, ``l_reconstruction``, and optionally ``n_ever_active`` (features with a non-zero code on
    at least one row) and ``frac_dead``."""
    x = batch.to(ensemble.device, non_blocking=True).float()
    losses, aux = ensemble.forward_batch(x)
    total_var = (x - x.mean(dim=0)).pow(2).mean()
    out = {
        "l_reconstruction": losses.get("l_reconstruction", losses["loss"]),
        "mean_l0": aux["c"].count_nonzero(dim=-1).float().mean(dim=-1),
    }
    out["fvu"] = out["l_reconstruction"] / total_var
    if n_ever_active:
        active = (aux["c"].dense() != 0).any(dim=1)           # [M, n]
        out["n_ever_active"] = active.sum(dim=-1)
        out["frac_dead"] = 1.0 - active.float().mean(dim=-1)
    return out
