"""world_size-2 gloo tests of the multi-GPU host logic (model-axis sharding, end-of-chunk metric gather, export
gather). Runs on CPU; rendezvous on 127.0.0.1."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sparse_coding_b200.sharding import gather_learned_dicts, gather_metrics, shard_models, shard_slices


def test_shard_slices():
    assert shard_slices(128, 8) == [(16 * r, 16 * r + 16) for r in range(8)]
    assert shard_slices(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard_slices(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    for n, w in ((128, 8), (10, 4), (7, 3)):
        sl = shard_slices(n, w)
        assert sl[0][0] == 0 and sl[-1][1] == n and all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
    assert shard_models(list(range(10)), 1, 4) == [3, 4, 5]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_models, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sparse_coding_b200 as S
    from sparse_coding_b200.train_loop import unstacked_to_learned_dicts
    torch.manual_seed(0)
    l1s = [10 ** (-4 + 0.5 * i) for i in range(n_models)]
    models = [S.FunctionalTiedSAE.init(8, 16, a) for a in l1s]            # every rank builds the same sweep ...
    sizes = [hi - lo for lo, hi in shard_slices(n_models, world)]
    mine = shard_models(models, rank, world)                               # ... and keeps its shard
    ens = S.FunctionalEnsemble(mine, S.FunctionalTiedSAE, S.adam, {"lr": 1e-3}, device="cpu")
    # stand-in for the per-model losses of the shard (the engine itself needs a GPU): l1_alpha and a rank tag
    local = torch.stack([ens.buffers["l1_alpha"], torch.full((len(mine),), float(rank))], dim=1)
    allm = gather_metrics(local, sizes)
    assert allm.shape == (n_models, 2)
    torch.testing.assert_close(allm[:, 0], torch.tensor(l1s))
    expect_rank = torch.cat([torch.full((s,), float(r)) for r, s in enumerate(sizes)])
    assert torch.equal(allm[:, 1], expect_rank)
    dicts = unstacked_to_learned_dicts(ens, {"dict_size": 16}, ["dict_size"], ["l1_alpha"])
    merged = gather_learned_dicts(dicts, dst=0)
    if rank == 0:
        assert len(merged) == n_models
        assert [round(h["l1_alpha"], 9) for _, h in merged] == [round(a, 9) for a in l1s]
        torch.save(merged, os.path.join(out_dir, "learned_dicts.pt"))
    else:
        assert merged is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_models", [4, 5])
def test_gloo_world2_shard_and_gather(tmp_path, n_models):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_models, str(tmp_path)), nprocs=2, join=True)
    loaded = torch.load(tmp_path / "learned_dicts.pt", weights_only=False)
    assert len(loaded) == n_models and loaded[0][0].encoder.shape == (16, 8)
