"""Generate tests/golden/*.pt by running the REFERENCE's own loss functions (HoagyC/sparse_coding @ 69c5ae0).

TEST INFRASTRUCTURE. Run in the build container only (needs /root/reference, which does not exist on the GPU
box):   python oracle/make_golden.py

The reference's ``autoencoders`` package imports three modules that are not installed here (torchopt, optree,
torchtyping). None of them is used by the loss functions themselves, so they are replaced by inert stubs; the
arithmetic that is recorded is 100 % the reference's (torch ops in autoencoders/sae_ensemble.py,
autoencoders/topk_encoder.py, autoencoders/learned_dict.py), driven the way FunctionalEnsemble.init_functions
drives it (ensemble.py:99-123): ``torch.vmap(torch.func.grad(sig.loss, has_aux=True))`` over stacked models, or a
per-model loop for TopK (``no_stacking=True``).

Every fixture stores its inputs (params, buffers, batch) and the reference outputs (loss_data, code, grads), so
tests can replay them against oracle/sae_oracle.py and against the CUDA engine without the reference tree.
"""
import os
import sys
import types

import torch

REF = os.environ.get("SCE_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    for name in ("torchopt", "optree"):
        sys.modules.setdefault(name, types.ModuleType(name))
    tt = types.ModuleType("torchtyping")

    class _TT:
        def __class_getitem__(cls, item):
            return cls

    tt.TensorType = _TT
    sys.modules.setdefault("torchtyping", tt)
    sys.path.insert(0, REF)
    import autoencoders.sae_ensemble as sae  # noqa
    import autoencoders.topk_encoder as topk  # noqa
    import autoencoders.learned_dict as ld  # noqa
    return sae, topk, ld


def stack(dicts):
    return {k: torch.stack([d[k] for d in dicts]) for k in dicts[0]}


def run_stacked(sig, models, batch):
    params, buffers = stack([m[0] for m in models]), stack([m[1] for m in models])
    f = torch.vmap(torch.func.grad(sig.loss, has_aux=True))
    with torch.no_grad():
        grads, (loss_data, aux) = f(params, buffers, batch.expand(len(models), *batch.shape))
    return params, buffers, grads, loss_data, aux


def run_looped(sig, models, batch):
    g = torch.func.grad(sig.loss, has_aux=True)
    outs = []
    with torch.no_grad():
        for p, b in models:
            outs.append(g(p, b, batch))
    grads = stack([o[0] for o in outs])
    loss_data = stack([o[1][0] for o in outs])
    aux = stack([o[1][1] for o in outs])
    return stack([m[0] for m in models]), stack([m[1] for m in models]), grads, loss_data, aux


def sparse_mix(B, d, n_feat, n_active, gen):
    """A sparse mixture of unit features + noise (the distribution of sc_datasets/random_dataset.py:76-142)."""
    feats = torch.randn(n_feat, d, generator=gen)
    feats = feats / feats.norm(dim=-1, keepdim=True)
    codes = torch.zeros(B, n_feat)
    for r in range(B):
        idx = torch.randperm(n_feat, generator=gen)[:n_active]
        codes[r, idx] = torch.rand(n_active, generator=gen)
    return codes @ feats + 0.05 * torch.randn(B, d, generator=gen)


def main():
    os.makedirs(OUT, exist_ok=True)
    sae, topk, ld = import_reference()
    torch.set_grad_enabled(False)
    fixtures = {}

    # ---- tied (FunctionalTiedSAE) -------------------------------------------------------------------------
    def tied_case(name, M, d, n, B, l1s, seed, data="gauss", dtype=torch.float32, bias_scale=0.0):
        torch.manual_seed(seed)
        gen = torch.Generator().manual_seed(seed + 1)
        models = []
        for l1 in l1s:
            p, b = sae.FunctionalTiedSAE.init(d, n, l1, dtype=dtype)
            b["bias_decay"] = torch.tensor(0.0, dtype=dtype)  # reference quirk Q1: init never creates it
            if bias_scale:
                p["encoder_bias"] = bias_scale * torch.randn(n, generator=gen).to(dtype)
            models.append((p, b))
        X = (torch.randn(B, d, generator=gen) if data == "gauss" else sparse_mix(B, d, 2 * n, 5, gen)).to(dtype)
        params, buffers, grads, loss_data, aux = run_stacked(sae.FunctionalTiedSAE, models, X)
        fixtures[name] = dict(kind="tied", params=params, buffers=buffers, batch=X, grads=grads,
                              loss_data=loss_data, c=aux["c"])

    tied_case("tied_small", 3, 32, 64, 48, [1e-3, 3e-3, 1e-2], 0)
    tied_case("tied_bias", 2, 64, 128, 96, [1e-4, 1e-2], 1, data="mix", bias_scale=0.05)
    tied_case("tied_f64", 2, 32, 64, 40, [1e-3, 1e-2], 2, dtype=torch.float64, bias_scale=0.02)

    # ---- tied with non-trivial centring -------------------------------------------------------------------
    torch.manual_seed(3)
    gen = torch.Generator().manual_seed(4)
    d, n, B = 32, 64, 40
    rot, _ = torch.linalg.qr(torch.randn(d, d, generator=gen))
    trans = 0.3 * torch.randn(d, generator=gen)
    scale = 0.5 + torch.rand(d, generator=gen)
    models = []
    for l1 in (1e-3, 1e-2):
        p, b = sae.FunctionalTiedSAE.init(d, n, l1, translation=trans.clone(), rotation=rot.clone(),
                                          scaling=scale.clone())
        b["bias_decay"] = torch.tensor(0.0)
        models.append((p, b))
    X = torch.randn(B, d, generator=gen)
    params, buffers, grads, loss_data, aux = run_stacked(sae.FunctionalTiedSAE, models, X)
    fixtures["tied_centered"] = dict(kind="tied", params=params, buffers=buffers, batch=X, grads=grads,
                                     loss_data=loss_data, c=aux["c"])

    # ---- untied (FunctionalSAE) with bias decay -----------------------------------------------------------
    torch.manual_seed(5)
    gen = torch.Generator().manual_seed(6)
    d, n, B = 48, 96, 64
    models = []
    for l1, bd in ((1e-3, 0.0), (3e-3, 0.01), (1e-2, 0.1)):
        p, b = sae.FunctionalSAE.init(d, n, l1, bias_decay=bd)
        p["encoder_bias"] = 0.05 * torch.randn(n, generator=gen)
        models.append((p, b))
    X = torch.randn(B, d, generator=gen)
    params, buffers, grads, loss_data, aux = run_stacked(sae.FunctionalSAE, models, X)
    fixtures["untied_small"] = dict(kind="untied", params=params, buffers=buffers, batch=X, grads=grads,
                                    loss_data=loss_data, c=aux["c"])

    # ---- masked variants (different dict sizes in one stack) ----------------------------------------------
    torch.manual_seed(7)
    gen = torch.Generator().manual_seed(8)
    d, nmax, B = 32, 96, 48
    X = torch.randn(B, d, generator=gen)
    models = [sae.FunctionalMaskedTiedSAE.init(d, nd, nmax, l1) for nd, l1 in ((32, 1e-3), (64, 1e-3), (96, 1e-2))]
    params, buffers, grads, loss_data, aux = run_stacked(sae.FunctionalMaskedTiedSAE, models, X)
    fixtures["masked_tied"] = dict(kind="masked_tied", params=params, buffers=buffers, batch=X, grads=grads,
                                   loss_data=loss_data, c=aux["c"])
    models = [sae.FunctionalMaskedSAE.init(d, nd, nmax, l1) for nd, l1 in ((32, 1e-3), (64, 1e-3), (96, 1e-2))]
    params, buffers, grads, loss_data, aux = run_stacked(sae.FunctionalMaskedSAE, models, X)
    fixtures["masked_untied"] = dict(kind="masked_untied", params=params, buffers=buffers, batch=X, grads=grads,
                                     loss_data=loss_data, c=aux["c"])

    # ---- TopK (no_stacking loop, ensemble.py:102-116) -----------------------------------------------------
    torch.manual_seed(9)
    gen = torch.Generator().manual_seed(10)
    d, n, B = 32, 128, 40
    X = torch.randn(B, d, generator=gen)
    models = [topk.TopKEncoder.init(d, n, k) for k in (4, 8, 16)]
    params, buffers, grads, loss_data, aux = run_looped(topk.TopKEncoder, models, X)
    fixtures["topk_small"] = dict(kind="topk", params=params, buffers=buffers, batch=X, grads=grads,
                                  loss_data=loss_data, c=aux["c"])

    # ---- BASELINE config 1: one TiedSAE d=128 n=256 L1=1e-3 B=1024, Gaussian activations ------------------
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(0)
    p, b = sae.FunctionalTiedSAE.init(128, 256, 1e-3)
    b["bias_decay"] = torch.tensor(0.0)
    X = torch.randn(1024, 128, generator=gen)
    params, buffers, grads, loss_data, aux = run_stacked(sae.FunctionalTiedSAE, [(p, b)], X)
    fixtures["cfg1"] = dict(kind="tied", params=params, buffers=buffers, batch=X, grads=grads,
                            loss_data=loss_data, c_nnz=aux["c"].count_nonzero(dim=-1),
                            c_sum=aux["c"].double().sum(dim=-1))

    # ---- LearnedDict inference API (learned_dict.py:129-215, topk_encoder.py:49-62) -----------------------
    torch.manual_seed(11)
    gen = torch.Generator().manual_seed(12)
    d, n, B = 32, 64, 24
    X = torch.randn(B, d, generator=gen)
    enc = torch.randn(n, d, generator=gen)
    dec = torch.randn(n, d, generator=gen)
    bias = 0.1 * torch.randn(n, generator=gen)
    rot, _ = torch.linalg.qr(torch.randn(d, d, generator=gen))
    trans = 0.3 * torch.randn(d, generator=gen)
    scale = 0.5 + torch.rand(d, generator=gen)
    tied = ld.TiedSAE(enc, bias, centering=(trans, rot, scale), norm_encoder=True)
    untied = ld.UntiedSAE(enc, dec, bias)
    tk = topk.TopKLearnedDict(dec / dec.norm(dim=-1, keepdim=True), 6)
    fixtures["learned_dicts"] = dict(
        kind="learned_dicts", batch=X, encoder=enc, decoder=dec, bias=bias, rot=rot, trans=trans, scale=scale,
        tied_encode=tied.encode(tied.center(X)), tied_predict=tied.predict(X), tied_dict=tied.get_learned_dict(),
        untied_encode=untied.encode(X), untied_predict=untied.predict(X), untied_dict=untied.get_learned_dict(),
        topk_encode=tk.encode(X), topk_predict=tk.predict(X), topk_k=6)

    for name, fx in fixtures.items():
        path = os.path.join(OUT, name + ".pt")
        torch.save(fx, path)
        print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
