"""CPU oracle for the ensemble sparse-autoencoder training step.

*** TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT PATH. ***
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; ``sparse_coding_b200`` never does and fails loudly without its CUDA
extension.

It restates, in plain PyTorch (CPU, fp32 or fp64), the algorithm of the reference hot path
(HoagyC/sparse_coding @ 69c5ae0):

* the per-model losses        autoencoders/sae_ensemble.py:53-78   (FunctionalSAE.loss, untied)
                              autoencoders/sae_ensemble.py:135-162 (FunctionalTiedSAE.loss)
                              autoencoders/sae_ensemble.py:347-373, 418-444 (masked variants)
                              autoencoders/topk_encoder.py:19-40   (TopKEncoder.encode / loss)
* the stacked ensemble step   autoencoders/ensemble.py:175-193     (FunctionalEnsemble.step_batch)
* the optimiser               torchopt 0.7.1 ``adam`` (requirements.txt:132) — a THIRD-PARTY dependency that
                              is not vendored under /root/reference and not installable offline.

Pinning status (see DESIGN.md "Oracle"):
  - loss values, codes and parameter gradients: PINNED against the reference's own loss functions, imported from
    /root/reference with stub modules for its missing imports (oracle/make_golden.py writes tests/golden/*.pt;
    tests/test_oracle.py replays them without the reference tree).
  - Adam arithmetic: **parity unpinned** — torchopt is absent; the update below follows torchopt's published
    ``scale_by_adam`` (mu/nu exponential moments, bias correction by ``1 - beta**count``,
    ``mu_hat / (sqrt(nu_hat + eps_root) + eps)``, then ``-lr``). ``count_mode="frozen_t1"`` reproduces the
    reference's step_batch quirk (ensemble.py:185-189 clones before copying, so the incremented count is dropped
    and the bias correction is always that of step 1); ``"standard"`` advances it.

Two independent formulations are provided and cross-checked in tests: closed-form gradients (``*_grads``) and the
restated loss functions under ``torch.func.grad`` / ``torch.vmap`` (``RefPortEnsemble``) — the latter mirrors the
reference's op sequence and is what the CPU baseline times.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch

Tensor = torch.Tensor
NORM_FLOOR = 1e-8  # sae_ensemble.py:59,137 clamp floor on row norms


# --------------------------------------------------------------------------------------------------------------
# forward pieces
# --------------------------------------------------------------------------------------------------------------
def unit_rows(mat: Tensor, floor: Optional[float] = NORM_FLOOR) -> Tuple[Tensor, Tensor]:
    """Rows scaled to unit L2 norm; returns (normalised, norms).  sae_ensemble.py:58-59 / 136-137.
    ``floor=None`` is the TopK variant, which divides by the raw norm (topk_encoder.py:31)."""
    s = mat.norm(dim=-1)
    if floor is not None:
        s = s.clamp(min=floor)
    return mat / s[..., None], s


def center(X: Tensor, trans: Tensor, rot: Tensor, scale: Tensor) -> Tensor:
    """sae_ensemble.py:126-128: ((x - t) @ rot^T) * scale."""
    return ((X - trans[None, :]) @ rot.T) * scale[None, :]


def tied_forward(E, b, X, alpha, bias_decay=0.0, coef_mask=None) -> Dict[str, Tensor]:
    """FunctionalTiedSAE.loss forward on an already-centred batch (sae_ensemble.py:135-162).
    coef_mask (bool [n], True = unused coefficient) gives FunctionalMaskedTiedSAE (:347-373)."""
    W, s = unit_rows(E)
    Z = X @ W.T + b
    C = Z.clamp(min=0.0)
    if coef_mask is not None:
        C = C.masked_fill(coef_mask, 0.0)
    Xh = C @ W
    l_rec = (Xh - X).pow(2).mean()
    l_l1 = alpha * C.abs().sum(dim=-1).mean()
    l_bd = bias_decay * b.norm()
    return dict(W=W, s=s, Z=Z, c=C, x_hat=Xh, l_reconstruction=l_rec, l_l1=l_l1, l_bias_decay=l_bd,
                loss=l_rec + l_l1 + l_bd)


def untied_forward(E, b, D, X, alpha, bias_decay=0.0, coef_mask=None) -> Dict[str, Tensor]:
    """FunctionalSAE.loss forward (sae_ensemble.py:53-78); coef_mask -> FunctionalMaskedSAE (:418-444)."""
    Z = X @ E.T + b
    C = Z.clamp(min=0.0)
    if coef_mask is not None:
        C = C.masked_fill(coef_mask, 0.0)
    Dn, s = unit_rows(D)
    Xh = C @ Dn
    l_rec = (Xh - X).pow(2).mean()
    l_l1 = alpha * C.abs().sum(dim=-1).mean()
    l_bd = bias_decay * b.norm()
    return dict(W=Dn, s=s, Z=Z, c=C, x_hat=Xh, l_reconstruction=l_rec, l_l1=l_l1, l_bias_decay=l_bd,
                loss=l_rec + l_l1 + l_bd)


def topk_code(S: Tensor, k: int) -> Tuple[Tensor, Tensor]:
    """TopKEncoder.encode (topk_encoder.py:19-27): keep the k largest SIGNED scores per row, then ReLU.
    Returns (code, selected-mask)."""
    idx = torch.topk(S, k, dim=-1).indices
    sel = torch.zeros_like(S, dtype=torch.bool)
    sel.scatter_(-1, idx, True)
    code = torch.where(sel, S, torch.zeros_like(S)).clamp(min=0.0)
    return code, sel


def topk_forward(Dct, X, k) -> Dict[str, Tensor]:
    """TopKEncoder.loss forward (topk_encoder.py:29-40)."""
    Wn, s = unit_rows(Dct, floor=None)
    S = X @ Wn.T
    C, sel = topk_code(S, int(k))
    Xh = C @ Wn
    loss = (X - Xh).pow(2).mean()
    return dict(W=Wn, s=s, Z=S, c=C, sel=sel, x_hat=Xh, loss=loss)


# --------------------------------------------------------------------------------------------------------------
# closed-form gradients (cross-checked against autograd in tests/test_oracle.py)
# --------------------------------------------------------------------------------------------------------------
def _row_norm_jacobian(W, s, dW):
    """d/dE of W = E / max(||E||, floor): (dW - W <W, dW>) / s   (rows above the floor)."""
    return (dW - W * (W * dW).sum(-1, keepdim=True)) / s[:, None]


def _bias_decay_grad(b, bias_decay):
    nb = b.norm()
    if float(bias_decay) == 0.0 or float(nb) == 0.0:
        return torch.zeros_like(b)
    return bias_decay * b / nb


def tied_grads(E, b, X, alpha, bias_decay=0.0, coef_mask=None, active=None) -> Dict[str, Tensor]:
    """``active`` (bool [B, n], optional) overrides the ReLU activity pattern [z > 0]: the loss is discontinuous in
    its derivative where a pre-activation is within rounding of zero, so a checker comparing two precisions must
    be able to pin the pattern for those (measure-zero) coefficients."""
    f = tied_forward(E, b, X, alpha, bias_decay, coef_mask)
    B, d = X.shape
    G = 2.0 * (f["x_hat"] - X) / (B * d)                       # dL/dx_hat
    pos = (f["c"] > 0) if active is None else active
    dC = G @ f["W"].T + (alpha / B) * pos.to(X.dtype)           # sign(0) = 0 for the L1 term
    gate = (f["Z"] >= 0) if active is None else (active | (f["Z"] == 0))  # clamp passes gradient at exactly 0
    if coef_mask is not None:
        gate = gate & ~coef_mask
    dZ = dC * gate.to(X.dtype)
    db = dZ.sum(0) + _bias_decay_grad(b, bias_decay)
    dW = dZ.T @ X + f["c"].T @ G
    f.update(G=G, dZ=dZ, grads={"encoder": _row_norm_jacobian(f["W"], f["s"], dW), "encoder_bias": db})
    return f


def untied_grads(E, b, D, X, alpha, bias_decay=0.0, coef_mask=None, active=None) -> Dict[str, Tensor]:
    """``active``: as in :func:`tied_grads` (pins the ReLU activity pattern of near-kink coefficients)."""
    f = untied_forward(E, b, D, X, alpha, bias_decay, coef_mask)
    B, d = X.shape
    G = 2.0 * (f["x_hat"] - X) / (B * d)
    pos = (f["c"] > 0) if active is None else active
    dC = G @ f["W"].T + (alpha / B) * pos.to(X.dtype)
    gate = (f["Z"] >= 0) if active is None else (active | (f["Z"] == 0))
    if coef_mask is not None:
        gate = gate & ~coef_mask
    dZ = dC * gate.to(X.dtype)
    db = dZ.sum(0) + _bias_decay_grad(b, bias_decay)
    dE = dZ.T @ X
    dDn = f["c"].T @ G
    f.update(G=G, dZ=dZ, grads={"encoder": dE, "encoder_bias": db,
                                "decoder": _row_norm_jacobian(f["W"], f["s"], dDn)})
    return f


def topk_grads(Dct, X, k, support=None) -> Dict[str, Tensor]:
    """``support`` (bool [B, n], optional) replaces the oracle's own selection-and-positive pattern: torch.topk
    leaves ties unspecified and a k-th score within rounding of the (k+1)-th may be ranked differently
    by two precisions, so a checker comparing them pins the support and separately verifies that it is a valid top-k."""
    f = topk_forward(Dct, X, k)
    B, d = X.shape
    if support is not None:
        f["c"] = torch.where(support, f["Z"], torch.zeros_like(f["Z"])).clamp(min=0.0)
        f["sel"] = support
        f["x_hat"] = f["c"] @ f["W"]
        f["loss"] = (X - f["x_hat"]).pow(2).mean()
    G = 2.0 * (f["x_hat"] - X) / (B * d)
    dS = (G @ f["W"].T) * (f["sel"] & (f["Z"] > 0)).to(X.dtype)   # relu: zero gradient at exactly 0
    dW = dS.T @ X + f["c"].T @ G
    f.update(G=G, dZ=dS, grads={"dict": _row_norm_jacobian(f["W"], f["s"], dW)})
    return f


# --------------------------------------------------------------------------------------------------------------
# Adam (torchopt.adam restated; see module docstring for the pinning caveat)
# --------------------------------------------------------------------------------------------------------------
ADAM_DEFAULTS = dict(lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0)


def adam_update(p: Tensor, g: Tensor, mu: Tensor, nu: Tensor, t: int, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8,
                eps_root=0.0) -> None:
    """In-place: moments, bias correction with step number ``t`` (>= 1), parameter update."""
    mu.mul_(b1).add_(g, alpha=1.0 - b1)
    nu.mul_(b2).addcmul_(g, g, value=1.0 - b2)
    mu_hat = mu / (1.0 - b1 ** t)
    nu_hat = nu / (1.0 - b2 ** t)
    p.add_(mu_hat / ((nu_hat + eps_root).sqrt() + eps), alpha=-lr)


# --------------------------------------------------------------------------------------------------------------
# restated loss functions in DictSignature form (params, buffers, batch) -> (loss, (loss_data, aux))
# used under torch.func.grad + torch.vmap exactly as ensemble.py:99-123 does
# --------------------------------------------------------------------------------------------------------------
def sig_loss_tied(params, buffers, batch):
    Xc = center(batch, buffers["center_trans"], buffers["center_rot"], buffers["center_scale"])
    f = tied_forward(params["encoder"], params["encoder_bias"], Xc, buffers["l1_alpha"],
                     buffers["bias_decay"] if "bias_decay" in buffers else 0.0)
    data = {"loss": f["loss"], "l_reconstruction": f["l_reconstruction"], "l_l1": f["l_l1"]}
    return f["loss"], (data, {"c": f["c"]})


def sig_loss_untied(params, buffers, batch):
    f = untied_forward(params["encoder"], params["encoder_bias"], params["decoder"], batch, buffers["l1_alpha"],
                       buffers["bias_decay"])
    data = {"loss": f["loss"], "l_reconstruction": f["l_reconstruction"], "l_l1": f["l_l1"],
            "l_bias_decay": f["l_bias_decay"]}
    return f["loss"], (data, {"c": f["c"]})


def sig_loss_masked_tied(params, buffers, batch):
    f = tied_forward(params["encoder"], params["encoder_bias"], batch, buffers["l1_alpha"], 0.0,
                     buffers["coef_mask"])
    l = f["l_reconstruction"] + f["l_l1"]
    return l, ({"loss": l, "l_reconstruction": f["l_reconstruction"], "l_l1": f["l_l1"]}, {"c": f["c"]})


def sig_loss_masked_untied(params, buffers, batch):
    f = untied_forward(params["encoder"], params["encoder_bias"], params["decoder"], batch, buffers["l1_alpha"],
                       0.0, buffers["coef_mask"])
    l = f["l_reconstruction"] + f["l_l1"]
    return l, ({"loss": l, "l_reconstruction": f["l_reconstruction"], "l_l1": f["l_l1"]}, {"c": f["c"]})


def sig_loss_topk(params, buffers, batch):
    f = topk_forward(params["dict"], batch, int(buffers["sparsity"]))
    return f["loss"], ({"loss": f["loss"]}, {"c": f["c"]})


SIG_LOSSES = {
    "tied": sig_loss_tied,
    "untied": sig_loss_untied,
    "masked_tied": sig_loss_masked_tied,
    "masked_untied": sig_loss_masked_untied,
    "topk": sig_loss_topk,
}


def _stack(dicts: List[Dict[str, Tensor]]) -> Dict[str, Tensor]:
    return {k: torch.stack([d[k] for d in dicts]) for k in dicts[0]}


class RefPortEnsemble:
    """Restatement of FunctionalEnsemble (ensemble.py:68-193) without torchopt/optree: params and buffers of M
    models stacked on dim 0, ``vmap(grad(loss, has_aux=True))`` (or a per-model loop when ``no_stacking``), Adam,
    in-place apply. ``loss_fn`` is one of SIG_LOSSES or the reference's own ``sig.loss`` (make_golden.py)."""

    def __init__(self, models, loss_fn, lr=1e-3, count_mode="frozen_t1", no_stacking=False, **adam):
        params, buffers = zip(*models)
        self.n_models = len(models)
        self.params = _stack(list(params))
        self.buffers = _stack(list(buffers))
        self.loss_fn = loss_fn
        self.no_stacking = no_stacking
        self.hp = dict(ADAM_DEFAULTS)
        self.hp.update(adam)
        self.hp["lr"] = lr
        self.count_mode = count_mode
        self.t = 0
        self.mu = {k: torch.zeros_like(v) for k, v in self.params.items()}
        self.nu = {k: torch.zeros_like(v) for k, v in self.params.items()}
        g = torch.func.grad(loss_fn, has_aux=True)
        self._vgrad = g if no_stacking else torch.vmap(g)
        self._grad1 = g

    def grads(self, batch: Tensor, expand_dims=True):
        with torch.no_grad():
            mb = batch.expand(self.n_models, *batch.shape) if expand_dims else batch
            if not self.no_stacking:
                return self._vgrad(self.params, self.buffers, mb)
            gs, ls, auxs = [], [], []
            for i in range(self.n_models):
                p = {k: v[i] for k, v in self.params.items()}
                b = {k: v[i] for k, v in self.buffers.items()}
                g, (l, a) = self._grad1(p, b, mb[i])
                gs.append(g)
                ls.append(l)
                auxs.append(a)
            return _stack(gs), (_stack(ls), _stack(auxs))

    def step_batch(self, batch: Tensor, expand_dims=True):
        grads, (loss, aux) = self.grads(batch, expand_dims)
        with torch.no_grad():
            self.t += 1
            t = 1 if self.count_mode == "frozen_t1" else self.t
            for k in self.params:
                adam_update(self.params[k], grads[k], self.mu[k], self.nu[k], t, **self.hp)
        return loss, aux


# --------------------------------------------------------------------------------------------------------------
# metrics that define "FVU vs ref" (standard_metrics.py:305-314)
# --------------------------------------------------------------------------------------------------------------
def fvu(x: Tensor, x_hat: Tensor) -> Tensor:
    return (x - x_hat).pow(2).mean() / (x - x.mean(dim=0)).pow(2).mean()


def xavier_uniform_bound(n, d):
    return math.sqrt(6.0 / (n + d))
