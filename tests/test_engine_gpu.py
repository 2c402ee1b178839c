"""Parity of the CUDA engine (through the C ABI, via sparse_coding_b200.FunctionalEnsemble) against
 (a) the golden vectors recorded from the reference's own loss functions (tests/golden, oracle/make_golden.py) and
 (b) the oracle (oracle/sae_oracle.py) on seeded inputs, single steps and multi-step trajectories.

Tolerances (BASELINE.json north_star: "within 1e-4 rel on reconstructed activations and loss"):
  x̂, code : ||a - b||_2 / ||b||_2 <= 1e-4          losses : |a - b| / |b| <= 1e-4
  gradients: norm-relative <= 2e-4 (3-pass backward)   trajectories: see test docstrings
"""
import math

import pytest
import torch

from oracle import sae_oracle as O

pytestmark = pytest.mark.gpu

REL = 1e-4


def relnorm(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


def _sigs():
    import sparse_coding_b200 as S
    return {"tied": S.FunctionalTiedSAE, "untied": S.FunctionalSAE, "masked_tied": S.FunctionalMaskedTiedSAE,
            "masked_untied": S.FunctionalMaskedSAE, "topk": S.TopKEncoder}


def _models(fx):
    M = next(iter(fx["params"].values())).shape[0]
    return [({k: v[i].float().clone() for k, v in fx["params"].items()},
             {k: (v[i].float().clone() if v.dtype.is_floating_point else v[i].clone()) for k, v in fx["buffers"].items()})
            for i in range(M)]


def _ensemble(fx, **kw):
    import sparse_coding_b200 as S
    return S.FunctionalEnsemble(_models(fx), _sigs()[fx["kind"]], S.adam, {"lr": 1e-3}, device="cuda", **kw)


GOLDEN = ["tied_small", "tied_bias", "tied_f64", "tied_centered", "untied_small", "masked_tied", "masked_untied",
          "topk_small"]


@pytest.mark.parametrize("name", GOLDEN)
def test_golden_forward_and_grads(golden, name):
    """Engine vs the reference's recorded loss_data / code / gradients on identical params and batch."""
    fx = golden(name)
    ens = _ensemble(fx)
    X = fx["batch"].float().cuda()
    grads, (loss, aux) = ens.grads_batch(X)
    for k, ref in fx["loss_data"].items():
        got = loss[k].cpu().double()
        assert torch.allclose(got, ref.double(), rtol=REL, atol=1e-9), (name, k, got, ref)
    c = aux["c"].dense()
    assert relnorm(c, fx["c"]) <= REL
    # sparsity pattern: identical wherever the reference activation is not within rounding of zero
    ref_c = fx["c"].float()
    big = ref_c.abs() > 1e-4 * ref_c.abs().max()
    assert torch.equal((c.cpu() != 0)[big], (ref_c != 0)[big])
    nnz_ref = fx["c"].count_nonzero(dim=-1).float().mean(dim=-1)
    assert torch.allclose(aux["c"].count_nonzero(dim=-1).float().mean(dim=-1).cpu(), nnz_ref, rtol=2e-2, atol=0.51)
    for k, ref in fx["grads"].items():
        assert relnorm(grads[k], ref) <= 2e-4, (name, k, relnorm(grads[k], ref))


@pytest.mark.parametrize("name", ["tied_small", "tied_centered", "untied_small", "masked_tied", "topk_small"])
def test_golden_reconstruction(golden, name):
    """x̂ (centred space) against the oracle evaluated on the golden inputs."""
    fx = golden(name)
    ens = _ensemble(fx)
    X = fx["batch"].float().cuda()
    _, _, x_hat = ens.forward_batch(X, return_x_hat=True)
    for i, (p, b) in enumerate(_models(fx)):
        Xd = fx["batch"].double()
        pd = {k: v.double() for k, v in p.items()}
        if fx["kind"] == "tied":
            Xc = O.center(Xd, b["center_trans"].double(), b["center_rot"].double(), b["center_scale"].double())
            f = O.tied_forward(pd["encoder"], pd["encoder_bias"], Xc, float(b["l1_alpha"]))
        elif fx["kind"] == "untied":
            f = O.untied_forward(pd["encoder"], pd["encoder_bias"], pd["decoder"], Xd, float(b["l1_alpha"]))
        elif fx["kind"] == "masked_tied":
            f = O.tied_forward(pd["encoder"], pd["encoder_bias"], Xd, float(b["l1_alpha"]), 0.0, b["coef_mask"])
        else:
            f = O.topk_forward(pd["dict"], Xd, int(b["sparsity"]))
        assert relnorm(x_hat[i], f["x_hat"]) <= REL, (name, i, relnorm(x_hat[i], f["x_hat"]))


def test_cfg1_golden(golden):
    """BASELINE config 1 (d=128, n=256, B=1024, L1=1e-3): losses, per-row nnz and gradients of the reference."""
    fx = golden("cfg1")
    ens = _ensemble(fx)
    grads, (loss, aux) = ens.grads_batch(fx["batch"].cuda())
    for k, ref in fx["loss_data"].items():
        assert torch.allclose(loss[k].cpu(), ref, rtol=REL, atol=0), (k, loss[k], ref)
    c = aux["c"].dense()[0].cpu()
    assert relnorm(c.double().sum(-1), fx["c_sum"][0]) <= REL
    assert (c.count_nonzero(dim=-1) - fx["c_nnz"][0]).abs().max() <= 1   # a score within rounding of 0 may flip
    for k, ref in fx["grads"].items():
        assert relnorm(grads[k], ref) <= 2e-4


def _random_tied(M, d, n, seed, l1=(1e-4, 1e-2), bias=0.02):
    import sparse_coding_b200 as S
    torch.manual_seed(seed)
    models = []
    for a in torch.logspace(math.log10(l1[0]), math.log10(l1[1]), M).tolist():
        p, b = S.FunctionalTiedSAE.init(d, n, a)
        p["encoder_bias"] = bias * torch.randn(n)
        models.append((p, b))
    return models


@pytest.mark.parametrize("B", [1, 37, 128, 200, 300])
def test_ragged_batches_and_last_short_batch(B):
    """drop_last=False gives a short final batch: any B <= batch_max must be exact, including after
    a larger batch left stale rows in the workspace."""
    import sparse_coding_b200 as S
    models = _random_tied(2, 64, 192, 0)
    ens = S.FunctionalEnsemble([({k: v.clone() for k, v in p.items()}, b) for p, b in models], S.FunctionalTiedSAE,
                               S.adam, {"lr": 1e-3}, device="cuda")
    gen = torch.Generator().manual_seed(B)
    big = torch.randn(300, 64, generator=gen)
    ens.forward_batch(big.cuda())                     # fill the workspace with 300 rows first
    X = torch.randn(B, 64, generator=gen)
    grads, (loss, aux) = ens.grads_batch(X.cuda())
    code = aux["c"].dense().cpu()
    for i, (p, b) in enumerate(models):
        f = O.tied_forward(p["encoder"].double(), p["encoder_bias"].double(), X.double(), float(b["l1_alpha"]))
        # pre-activations within rounding of zero: take the engine's side of the kink (see tied_grads docstring)
        active = torch.where(f["Z"].abs() < 1e-5, code[i] > 0, f["Z"] > 0)
        f = O.tied_grads(p["encoder"].double(), p["encoder_bias"].double(), X.double(), float(b["l1_alpha"]),
                         active=active)
        assert abs(float(loss["loss"][i]) - float(f["loss"])) <= REL * abs(float(f["loss"]))
        assert relnorm(grads["encoder"][i], f["grads"]["encoder"]) <= 2e-4
        assert relnorm(grads["encoder_bias"][i], f["grads"]["encoder_bias"]) <= 2e-4
        assert relnorm(aux["c"].dense()[i], f["c"]) <= REL


def test_exact_zero_rows_clamp_gradient():
    """Q4: all-zero input rows with zero bias give z == 0 exactly; clamp(min=0) passes the gradient there."""
    import sparse_coding_b200 as S
    torch.manual_seed(0)
    p, b = S.FunctionalTiedSAE.init(32, 64, 1e-2)
    X = torch.randn(40, 32)
    X[5] = 0.0
    X[17] = 0.0
    ens = S.FunctionalEnsemble([({k: v.clone() for k, v in p.items()}, b)], S.FunctionalTiedSAE, S.adam, {"lr": 1e-3},
                               device="cuda")
    grads, (loss, aux) = ens.grads_batch(X.cuda())
    f = O.tied_grads(p["encoder"].double(), p["encoder_bias"].double(), X.double(), 1e-2)
    assert relnorm(grads["encoder"][0], f["grads"]["encoder"]) <= 2e-4
    assert relnorm(grads["encoder_bias"][0], f["grads"]["encoder_bias"]) <= 2e-4
    assert int(aux["c"].dense()[0, 5].count_nonzero()) == 0


@pytest.mark.parametrize("mode", ["frozen_t1", "standard"])
@pytest.mark.parametrize("kind", ["tied", "untied"])
def test_training_trajectory_matches_oracle(kind, mode):
    """30 optimiser steps, engine vs the restated reference step (RefPortEnsemble) from identical initial state on
    identical batches. Adam's update is sign-like where |g| is tiny, so parameters are compared in norm and the
    per-step losses to 1e-3."""
    import sparse_coding_b200 as S
    torch.manual_seed(1)
    d, n, B, M = 64, 256, 256, 3
    sig = S.FunctionalTiedSAE if kind == "tied" else S.FunctionalSAE
    models = []
    for a in (1e-4, 1e-3, 1e-2):
        p, b = sig.init(d, n, a) if kind == "tied" else sig.init(d, n, a, bias_decay=0.01)
        models.append((p, b))
    clone = lambda ms: [({k: v.clone() for k, v in p.items()}, {k: v.clone() for k, v in b.items()}) for p, b in ms]
    ens = S.FunctionalEnsemble(clone(models), sig, S.adam, {"lr": 1e-3}, device="cuda", adam_count_mode=mode)
    ref = O.RefPortEnsemble(clone(models), O.SIG_LOSSES[kind], lr=1e-3, count_mode=mode)
    gen = torch.Generator().manual_seed(2)
    feats = torch.randn(512, d, generator=gen)
    feats /= feats.norm(dim=-1, keepdim=True)
    for step in range(30):
        codes = (torch.rand(B, 512, generator=gen) < 0.02).float() * torch.rand(B, 512, generator=gen)
        X = codes @ feats + 0.01 * torch.randn(B, d, generator=gen)
        loss, aux = ens.step_batch(X.cuda())
        rloss, raux = ref.step_batch(X)
        for k in rloss:
            assert torch.allclose(loss[k].cpu(), rloss[k], rtol=1e-3, atol=1e-7), (step, k, loss[k], rloss[k])
    for k in ref.params:
        assert relnorm(ens.params[k], ref.params[k]) <= 2e-3, (k, relnorm(ens.params[k], ref.params[k]))
    assert relnorm(ens.optim_states["mu"]["encoder"], ref.mu["encoder"]) <= 1e-3
    assert relnorm(ens.optim_states["nu"]["encoder"], ref.nu["encoder"]) <= 1e-3


def test_topk_trajectory_matches_oracle():
    import sparse_coding_b200 as S
    torch.manual_seed(3)
    d, n, B = 64, 256, 128
    models = [S.TopKEncoder.init(d, n, k) for k in (4, 8, 16)]
    clone = lambda ms: [({k: v.clone() for k, v in p.items()}, {k: v.clone() for k, v in b.items()}) for p, b in ms]
    ens = S.FunctionalEnsemble(clone(models), S.TopKEncoder, S.adam, {"lr": 1e-3}, device="cuda", no_stacking=True)
    ref = O.RefPortEnsemble(clone(models), O.SIG_LOSSES["topk"], lr=1e-3, no_stacking=True)
    gen = torch.Generator().manual_seed(4)
    for step in range(10):
        X = torch.randn(B, d, generator=gen)
        loss, aux = ens.step_batch(X.cuda())
        rloss, raux = ref.step_batch(X)
        assert torch.allclose(loss["loss"].cpu(), rloss["loss"], rtol=1e-3), (step, loss["loss"], rloss["loss"])
    assert relnorm(ens.params["dict"], ref.params["dict"]) <= 2e-3


def test_config2_shape_properties():
    """BASELINE config 2 at full size (M=16, d=512, n=4096, B=8192): the oracle cannot run this in seconds, so
    check x̂ / code on a row slice against fp64 torch on the GPU, the nnz counter against the dense code, loss
    consistency (loss == l_rec + l_l1) and that 3 steps reduce every model's loss."""
    import sparse_coding_b200 as S
    M, d, n, B = 16, 512, 4096, 8192
    torch.manual_seed(0)
    models = [S.FunctionalTiedSAE.init(d, n, float(a)) for a in torch.logspace(-4, -2, M)]
    ens = S.FunctionalEnsemble(models, S.FunctionalTiedSAE, S.adam, {"lr": 1e-3}, device="cuda")
    gen = torch.Generator().manual_seed(1)
    X = torch.randn(B, d, generator=gen).cuda()
    loss0, aux0, x_hat = ens.forward_batch(X, return_x_hat=True)
    rows = torch.arange(0, B, 97, device="cuda")
    for m in (0, 7, 15):
        E = ens.params["encoder"][m].double()
        f = O.tied_forward(E, ens.params["encoder_bias"][m].double(), X[rows].double(), float(ens.buffers["l1_alpha"][m]))
        assert relnorm(x_hat[m][rows], f["x_hat"]) <= REL
    c = aux0["c"].dense()
    assert torch.allclose(c.count_nonzero(dim=-1).float().mean(dim=-1), aux0["c"].count_nonzero(dim=-1).float().mean(dim=-1),
                          rtol=1e-6)
    l1_ref = ens.buffers["l1_alpha"] * c.sum(dim=-1).mean(dim=-1)
    assert torch.allclose(loss0["l_l1"], l1_ref, rtol=REL)
    rec_ref = (x_hat - X[None]).double().pow(2).mean(dim=(1, 2)).float()
    assert torch.allclose(loss0["l_reconstruction"], rec_ref, rtol=REL)
    assert torch.allclose(loss0["loss"], loss0["l_reconstruction"] + loss0["l_l1"], rtol=1e-6)
    del c, x_hat
    first = None
    for _ in range(3):
        loss, _ = ens.step_batch(X)
        first = loss["loss"].clone() if first is None else first
    assert bool((loss["loss"] < first).all())
    assert all(torch.isfinite(v).all() for v in ens.params.values())
