"""The oracle (oracle/sae_oracle.py) against the golden vectors recorded from the reference's own loss functions
(oracle/make_golden.py), and its two formulations (closed form vs autograd) against each other."""
import pytest
import torch

from oracle import sae_oracle as O

TIED = ["tied_small", "tied_bias", "tied_f64", "tied_centered", "cfg1"]


def _model(fx, i):
    return {k: v[i] for k, v in fx["params"].items()}, {k: v[i] for k, v in fx["buffers"].items()}


def _tol(dtype):
    return dict(rtol=1e-10, atol=1e-12) if dtype == torch.float64 else dict(rtol=2e-5, atol=2e-7)


def _closed_form(fx, i):
    p, b = _model(fx, i)
    X = fx["batch"]
    kind = fx["kind"]
    if kind == "tied":
        Xc = O.center(X, b["center_trans"], b["center_rot"], b["center_scale"])
        return O.tied_grads(p["encoder"], p["encoder_bias"], Xc, b["l1_alpha"], b["bias_decay"])
    if kind == "untied":
        return O.untied_grads(p["encoder"], p["encoder_bias"], p["decoder"], X, b["l1_alpha"], b["bias_decay"])
    if kind == "masked_tied":
        return O.tied_grads(p["encoder"], p["encoder_bias"], X, b["l1_alpha"], 0.0, b["coef_mask"])
    if kind == "masked_untied":
        return O.untied_grads(p["encoder"], p["encoder_bias"], p["decoder"], X, b["l1_alpha"], 0.0, b["coef_mask"])
    if kind == "topk":
        return O.topk_grads(p["dict"], X, int(b["sparsity"]))
    raise AssertionError(kind)


@pytest.mark.parametrize("name", TIED + ["untied_small", "masked_tied", "masked_untied", "topk_small"])
def test_closed_form_matches_reference(golden, name):
    fx = golden(name)
    M = next(iter(fx["params"].values())).shape[0]
    for i in range(M):
        f = _closed_form(fx, i)
        tol = _tol(fx["batch"].dtype)
        for k, ref in fx["loss_data"].items():
            torch.testing.assert_close(f[k], ref[i], **tol)
        if "c" in fx:
            torch.testing.assert_close(f["c"], fx["c"][i], **tol)
            assert torch.equal(f["c"] != 0, fx["c"][i] != 0)
        else:
            assert torch.equal(f["c"].count_nonzero(dim=-1), fx["c_nnz"][i])
        for k, ref in fx["grads"].items():
            torch.testing.assert_close(f["grads"][k], ref[i], **tol)


@pytest.mark.parametrize("name", ["tied_small", "tied_centered", "untied_small", "masked_tied", "masked_untied"])
def test_port_vmap_grad_matches_reference(golden, name):
    fx = golden(name)
    M = next(iter(fx["params"].values())).shape[0]
    models = [_model(fx, i) for i in range(M)]
    ens = O.RefPortEnsemble(models, O.SIG_LOSSES[fx["kind"]])
    grads, (loss, aux) = ens.grads(fx["batch"])
    tol = _tol(fx["batch"].dtype)
    for k, ref in fx["loss_data"].items():
        torch.testing.assert_close(loss[k], ref, **tol)
    torch.testing.assert_close(aux["c"], fx["c"], **tol)
    for k, ref in fx["grads"].items():
        torch.testing.assert_close(grads[k], ref, **tol)


def test_port_topk_loop_matches_reference(golden):
    fx = golden("topk_small")
    models = [_model(fx, i) for i in range(3)]
    ens = O.RefPortEnsemble(models, O.SIG_LOSSES["topk"], no_stacking=True)
    grads, (loss, aux) = ens.grads(fx["batch"])
    torch.testing.assert_close(loss["loss"], fx["loss_data"]["loss"], rtol=2e-5, atol=2e-7)
    torch.testing.assert_close(aux["c"], fx["c"], rtol=2e-5, atol=2e-7)
    torch.testing.assert_close(grads["dict"], fx["grads"]["dict"], rtol=2e-5, atol=2e-7)
    # Q8: at most k non-zeros per row, fewer when a selected score is negative
    for i, k in enumerate((4, 8, 16)):
        assert int(fx["c"][i].count_nonzero(dim=-1).max()) <= k


def test_clamp_passes_gradient_at_exact_zero():
    """Q4: clamp(min=0) has gradient 1 at z == 0 (all-zero input row, zero bias), relu would give 0."""
    torch.manual_seed(0)
    E = torch.randn(8, 4, dtype=torch.float64)
    b = torch.zeros(8, dtype=torch.float64)
    X = torch.randn(5, 4, dtype=torch.float64)
    X[2] = 0.0
    cf = O.tied_grads(E, b, X, 1e-2)
    p = {"encoder": E, "encoder_bias": b}
    bu = {"center_trans": torch.zeros(4, dtype=torch.float64), "center_rot": torch.eye(4, dtype=torch.float64),
          "center_scale": torch.ones(4, dtype=torch.float64), "l1_alpha": torch.tensor(1e-2, dtype=torch.float64),
          "bias_decay": torch.tensor(0.0, dtype=torch.float64)}
    g, _ = torch.func.grad(O.sig_loss_tied, has_aux=True)(p, bu, X)
    torch.testing.assert_close(cf["grads"]["encoder"], g["encoder"], rtol=1e-12, atol=1e-14)
    torch.testing.assert_close(cf["grads"]["encoder_bias"], g["encoder_bias"], rtol=1e-12, atol=1e-14)
    assert (cf["Z"][2] == 0).all() and (cf["dZ"][2] == 0).all()  # x = x_hat = 0 there, so the passed gradient is 0


def test_adam_modes():
    torch.manual_seed(0)
    p0 = torch.randn(6, 3)
    g = torch.randn(6, 3)
    # frozen_t1: first step is -lr * sign(g) (up to eps), and stays bias-corrected as step 1
    p, mu, nu = p0.clone(), torch.zeros(6, 3), torch.zeros(6, 3)
    O.adam_update(p, g, mu, nu, 1, lr=1e-3)
    torch.testing.assert_close(p, p0 - 1e-3 * g.sign(), rtol=0, atol=1e-7)
    # standard mode equals torch.optim.Adam
    q = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([q], lr=1e-3)
    p, mu, nu = p0.clone(), torch.zeros(6, 3), torch.zeros(6, 3)
    for t in range(1, 6):
        gt = g * t
        q.grad = gt.clone()
        opt.step()
        O.adam_update(p, gt, mu, nu, t, lr=1e-3)
    torch.testing.assert_close(p, q.detach(), rtol=1e-5, atol=1e-7)


def test_fvu_formula():
    x = torch.randn(50, 7)
    assert float(O.fvu(x, x)) == 0.0
    torch.testing.assert_close(O.fvu(x, x.mean(0).expand_as(x)), torch.tensor(1.0))
