"""On-device evaluation of a training ensemble: the quantities the reference computes per exported
dictionary on the CPU — FVU (standard_metrics.py:310-314), mean L0 and per-feature activation frequency (:305-308),
features ever active (:441-454) — obtained for all M models at once from the engine's forward pass: no parameter
update and NO dense fp32 code. FVU and L0 come from the fused loss / nnz counters of the GEMM epilogues; the
per-feature activation counts are column sums of the [c > 0] activity-mask plane that the encode epilogue (or the top-k
selection) writes for the backward pass (libsce ``sce_active_counts``), accumulated over as many batches as the
held-out set has.

FVU is taken in the space the model reconstructs (the centred space for FunctionalTiedSAE with a non-trivial
centring: an orthogonal rotation leaves it unchanged, a non-uniform ``center_scale`` does not)."""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import torch

EVER_ACTIVE_THRESHOLD = 10   # standard_metrics.py:446: a feature counts as "ever active" above this many rows


def evaluate(ensemble, batch: torch.Tensor, n_ever_active: bool = False, threshold: int = 0,
             counts: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """One held-out batch [B, d] (CUDA or pinned host). Returns per-model tensors on the ensemble's device:
    ``fvu``, ``mean_l0``, ``l_reconstruction``, and with ``n_ever_active`` also ``feature_counts`` ([M, n] rows on
    which each feature fired; accumulated into ``counts`` when given), ``n_ever_active`` (features that fired on
    more than ``threshold`` rows) and ``frac_dead``."""
    x = batch.to(ensemble.device, non_blocking=True).float()
    losses, aux = ensemble.forward_batch(x)
    total_var = (x - x.mean(dim=0)).pow(2).mean()
    out = {
        "l_reconstruction": losses.get("l_reconstruction", losses["loss"]),
        "mean_l0": aux["c"].count_nonzero(dim=-1).float().mean(dim=-1),
    }
    out["fvu"] = out["l_reconstruction"] / total_var
    if n_ever_active:
        counts = ensemble.active_counts(x.shape[-2], counts)
        out["feature_counts"] = counts
        out["n_ever_active"] = (counts > threshold).sum(dim=-1)
        out["frac_dead"] = 1.0 - out["n_ever_active"].float() / counts.shape[-1]
    return out


def evaluate_batches(ensemble, batches: Iterable[torch.Tensor],
                     threshold: int = EVER_ACTIVE_THRESHOLD) -> Dict[str, torch.Tensor]:
    """A held-out set streamed through in batches (``batched_calc_feature_n_ever_active``, standard_metrics.py:446-454,
    for every model of the ensemble at once, plus FVU and L0 of the whole set): FVU = sum of squared residuals / total
    variance about the set's column means — exactly the reference's formula on the concatenated set —, row-weighted
    mean L0, per-feature activation counts and frequencies, features active on more than ``threshold`` rows."""
    dev = torch.device(ensemble.device)
    sq = torch.zeros(ensemble.n_models, dtype=torch.float64, device=dev)
    l0 = torch.zeros(ensemble.n_models, dtype=torch.float64, device=dev)
    s1 = s2 = None
    rows, counts = 0, None
    for b in batches:
        x = b.to(dev, non_blocking=True).float()
        B, d = x.shape
        losses, aux = ensemble.forward_batch(x)
        sq += losses.get("l_reconstruction", losses["loss"]).double() * (B * d)
        l0 += aux["c"].count_nonzero(dim=-1).float().mean(dim=-1).double() * B
        counts = ensemble.active_counts(B, counts)
        xd = x.double()
        s1 = xd.sum(0) if s1 is None else s1 + xd.sum(0)
        s2 = xd.pow(2).sum(0) if s2 is None else s2 + xd.pow(2).sum(0)
        rows += B
    if rows == 0:
        raise ValueError("evaluate_batches needs at least one batch")
    total = (s2 - s1 * s1 / rows).sum()                      # sum over elements of (x - column mean)^2
    n_act = (counts > threshold).sum(dim=-1)
    return {"fvu": (sq / total).float(), "mean_l0": (l0 / rows).float(), "feature_counts": counts,
            "feature_frequency": counts.float() / rows, "n_ever_active": n_act,
            "frac_dead": 1.0 - n_act.float() / counts.shape[-1], "rows": rows}
