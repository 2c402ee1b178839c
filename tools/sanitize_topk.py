"""Small top-k runs for compute-sanitizer (memcheck / racecheck): sparse path (d % 32 == 0), dense path (d % 32 != 0),
the radix fall-back of the selection (all-equal rows, k > 256), the chunk-maxima path, forward with x_hat, a short last
batch; and the centring kernels of the tied variant."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sparse_coding_b200 as S

torch.manual_seed(0)
for d, n, ks, B in ((64, 256, (4, 9, 16), 96), (40, 264, (3, 8), 70), (64, 1200, (300, 5), 33),
                    (64, 4096, (16, 40), 64)):      # 128 chunk maxima per row: the fused selection path
    models = [S.TopKEncoder.init(d, n, k) for k in ks]
    ens = S.FunctionalEnsemble(models, S.TopKEncoder, S.adam, {"lr": 1e-3}, device="cuda", no_stacking=True)
    X = torch.randn(B, d).cuda()
    X[3] = 0.0                                   # all scores equal: every key is a candidate -> radix fall-back
    for _ in range(2):
        loss, aux = ens.step_batch(X)
    loss, aux, xh = ens.forward_batch(X[: B - 7], return_x_hat=True)
    c = aux["c"].dense()
    torch.cuda.synchronize()
    print(d, n, ks, "loss", [round(float(v), 5) for v in loss["loss"]], "nnz", [float(v) for v in (c != 0).sum(-1).float().mean(-1)])
# device-side centring (shared and per-model batches)
d, n, B = 64, 128, 50
models = []
for i in range(2):
    q, _ = torch.linalg.qr(torch.randn(d, d))
    models.append(S.FunctionalTiedSAE.init(d, n, 1e-3, translation=0.1 * torch.randn(d), rotation=q.contiguous(),
                                           scaling=0.5 + torch.rand(d)))
ens = S.FunctionalEnsemble(models, S.FunctionalTiedSAE, S.adam, {"lr": 1e-3}, device="cuda")
for X, ed in ((torch.randn(B, d).cuda(), True), (torch.randn(2, B, d).cuda(), False)):
    for _ in range(2):
        loss, aux = ens.step_batch(X, expand_dims=ed)
    torch.cuda.synchronize()
    print("centred", ed, [round(float(v), 5) for v in loss["loss"]])
print("ok")
