"""Soak / sanity run at BASELINE config-2 scale: train 16 tied SAEs (d=512, n=4096, L1 = logspace(-4,-2,16),
batch 8192, lr 1e-3) for N steps on a sparse mixture of 2048 ground-truth unit features
(sc_datasets/random_dataset.py semantics) and report, per model, FVU, mean L0, dead fraction and the mean max cosine
similarity (MMCS, standard_metrics.py:270-297) between ground-truth features and learned dictionary rows."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sparse_coding_b200 as S
from sparse_coding_b200.metrics import evaluate

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
M, d, n, B, n_gt = 16, 512, 4096, 8192, 2048
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(0)
feats = torch.randn(n_gt, d, device=dev, generator=gen)
feats /= feats.norm(dim=-1, keepdim=True)
probs = 0.99 ** torch.arange(n_gt, device=dev).float()
probs = probs / probs.sum() * 20.0                         # about 20 active features per row, decaying frequency


def batch(rows):
    active = torch.rand(rows, n_gt, device=dev, generator=gen) < probs
    codes = active.float() * torch.rand(rows, n_gt, device=dev, generator=gen)
    return codes @ feats + 0.01 * torch.randn(rows, d, device=dev, generator=gen)


torch.manual_seed(0)
models = [S.FunctionalTiedSAE.init(d, n, float(a)) for a in np.logspace(-4, -2, M)]
ens = S.FunctionalEnsemble(models, S.FunctionalTiedSAE, S.adam, {"lr": 1e-3}, device=dev)
held = batch(B)
first = evaluate(ens, held)
torch.cuda.synchronize()
t0 = time.perf_counter()
for s in range(steps):
    losses, aux = ens.step_batch(batch(B))
torch.cuda.synchronize()
dt = time.perf_counter() - t0
ev = evaluate(ens, held, n_ever_active=True)
W = ens.params["encoder"] / ens.params["encoder"].norm(dim=-1, keepdim=True)           # [M, n, d]
mmcs = torch.stack([(feats[:512] @ W[m].T).max(dim=-1).values.mean() for m in range(M)])  # 512 most frequent features
finite = all(torch.isfinite(v).all().item() for v in ens.params.values())
print(json.dumps({
    "steps": steps, "seconds_incl_datagen": dt, "all_finite": finite,
    "l1_alpha": [float(a) for a in ens.buffers["l1_alpha"]],
    "fvu_before": [round(float(v), 4) for v in first["fvu"]], "fvu_after": [round(float(v), 4) for v in ev["fvu"]],
    "mean_l0_after": [round(float(v), 1) for v in ev["mean_l0"]],
    "frac_dead_after": [round(float(v), 3) for v in ev["frac_dead"]],
    "mmcs_top512_gt_features": [round(float(v), 3) for v in mmcs]}))
