"""End-to-end chunk training: synthetic fp16 `{i}.pt` chunks on disk -> ChunkStreamer (pinned
staging + side-stream H2D, overlapped with training) -> device-side batch gather -> step -> export
`learned_dicts.pt`. Prints one JSON line with activations/s over the whole run (disk, copies, export included)."""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sparse_coding_b200 as S
from sparse_coding_b200.train_loop import train_on_chunks

M, d, n, B = 16, 512, 4096, 8192
rows, n_chunks = 1 << 19, 4            # 4 chunks x 512 Ki rows x 512 x fp16 = 0.5 GiB each
tmp = tempfile.mkdtemp()
data, out = os.path.join(tmp, "data"), os.path.join(tmp, "out")
os.makedirs(data)
gen = torch.Generator().manual_seed(0)
for i in range(n_chunks):
    torch.save(torch.randn(rows, d, generator=gen).half(), os.path.join(data, f"{i}.pt"))
torch.manual_seed(0)
models = [S.FunctionalTiedSAE.init(d, n, float(a)) for a in np.logspace(-4, -2, M)]
ens = S.FunctionalEnsemble(models, S.FunctionalTiedSAE, S.adam, {"lr": 1e-3}, device="cuda")
ens.step_batch(torch.randn(B, d).cuda())                       # plan + workspace
torch.cuda.synchronize()
t0 = time.perf_counter()
dicts = train_on_chunks(ens, {"device": "cuda", "dict_size": n}, data, out, B, ["dict_size"], ["l1_alpha"],
                        chunk_order=list(range(n_chunks)), center_activations=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
steps = n_chunks * (rows // B)
print(json.dumps({"what": "train_on_chunks end to end (disk -> pinned -> HBM -> gather -> step -> export)",
                  "activations_per_s": n_chunks * rows / dt, "seconds": dt, "steps": steps, "ms_per_step": dt / steps * 1e3,
                  "chunks": n_chunks, "rows_per_chunk": rows, "exported": len(dicts),
                  "checkpoint": os.path.exists(os.path.join(out, f"_{n_chunks - 1}", "learned_dicts.pt"))}))
