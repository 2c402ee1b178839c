"""Diagnostic (not a test): error of the engine's losses / code / gradients vs the fp64 oracle, next to the error of
the fp32 PyTorch restatement (the reference's own arithmetic) vs the same fp64 oracle — i.e. the noise floor."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparse_coding_b200 as S
from oracle import sae_oracle as O


def rn(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


def run(M, d, n, B, seed=0, bias=0.02, passes=(3, 3), prefill=0):
    torch.manual_seed(seed)
    models = []
    for a in torch.logspace(-4, -2, M).tolist():
        p, b = S.FunctionalTiedSAE.init(d, n, a)
        p["encoder_bias"] = bias * torch.randn(n)
        models.append((p, b))
    ens = S.FunctionalEnsemble([({k: v.clone() for k, v in p.items()}, b) for p, b in models], S.FunctionalTiedSAE,
                               S.adam, {"lr": 1e-3}, device="cuda", fwd_passes=passes[0], bwd_passes=passes[1])
    gen = torch.Generator().manual_seed(seed + 100)
    if prefill:
        ens.forward_batch(torch.randn(prefill, d, generator=gen).cuda())
    X = torch.randn(B, d, generator=gen)
    grads, (loss, aux) = ens.grads_batch(X.cuda())
    _, _, xh = ens.forward_batch(X.cuda(), return_x_hat=True)
    c = aux["c"]
    for i, (p, b) in enumerate(models):
        a = float(b["l1_alpha"])
        f64 = O.tied_grads(p["encoder"].double().cuda(), p["encoder_bias"].double().cuda(), X.double().cuda(), a)
        f32 = O.tied_grads(p["encoder"].cuda(), p["encoder_bias"].cuda(), X.cuda(), a)
        print(f"M{i} d={d} n={n} B={B} passes={passes} prefill={prefill} | "
              f"loss eng {abs(float(loss['loss'][i]) - float(f64['loss'])) / float(f64['loss']):.1e} "
              f"fp32 {abs(float(f32['loss']) - float(f64['loss'])) / float(f64['loss']):.1e} | "
              f"xhat eng {rn(xh[i], f64['x_hat']):.1e} fp32 {rn(f32['x_hat'], f64['x_hat']):.1e} | "
              f"dE eng {rn(grads['encoder'][i], f64['grads']['encoder']):.1e} fp32 {rn(f32['grads']['encoder'], f64['grads']['encoder']):.1e} | "
              f"db eng {rn(grads['encoder_bias'][i], f64['grads']['encoder_bias']):.1e} fp32 {rn(f32['grads']['encoder_bias'], f64['grads']['encoder_bias']):.1e}")


if __name__ == "__main__":
    for B in (1, 37, 64, 96, 128, 160, 200, 256, 300):
        run(2, 64, 192, B, prefill=300)
    run(2, 64, 192, 128, prefill=0)
    run(2, 128, 256, 1024)
    run(2, 512, 4096, 2048)
    run(2, 512, 4096, 2048, passes=(3, 1))
    run(2, 512, 4096, 2048, passes=(1, 1))
