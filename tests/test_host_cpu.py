"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol include/sce.h declares, argument
validation across the ABI, the host-side mirror of the reference interface (stacking, state_dict, optimiser
resolution, export types, checkpoint pickle names, wandb key format), and the loud failure when asked to compute
without CUDA."""
import ctypes as C
import io
import os
import pickle
import re

import pytest
import torch

import sparse_coding_b200 as S
from sparse_coding_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "sce.h")).read()
    declared = set(re.findall(r"\b(sce_[a-z_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.sce_version() == 201


def test_abi_validation_without_device():
    lib = _lib.load()
    desc = _lib.SceDesc(variant=0, n_models=2, d=64, n=128, batch_max=256, x_per_model=0, lr=1e-3, beta1=0.9,
                        beta2=0.999, eps=1e-8, eps_root=0.0, adam_count_mode=0, fwd_passes=3, bwd_passes=3,
                        norm_floor=1e-8)
    need = lib.sce_workspace_bytes(C.byref(desc))
    assert need > 0 and need % 1024 == 0
    desc.batch_max = 512
    assert lib.sce_workspace_bytes(C.byref(desc)) > need          # grows with the batch
    desc.d = 63                                                      # not a multiple of 8
    assert lib.sce_workspace_bytes(C.byref(desc)) == 0
    assert b"multiples of 8" in lib.sce_last_error()
    # operand arithmetic (sce_arith): both carry 4 bytes per operand element, so the workspace does not depend on it;
    # f16f8 needs 16-byte row pitches in its 8-bit planes, AUTO falls back to bf16x3 where that fails
    desc.d = 64
    sizes = []
    for arith in (_lib.SCE_ARITH_AUTO, _lib.SCE_ARITH_BF16X3, _lib.SCE_ARITH_F16F8):
        desc.arith = arith
        sizes.append(lib.sce_workspace_bytes(C.byref(desc)))
    assert sizes[0] > 0 and max(sizes) - min(sizes) <= 64 * 1024, sizes
    desc.d, desc.arith = 72, _lib.SCE_ARITH_F16F8                    # multiple of 8 but not of 16
    assert lib.sce_workspace_bytes(C.byref(desc)) == 0
    assert b"multiples of 16" in lib.sce_last_error()
    desc.arith = _lib.SCE_ARITH_AUTO
    assert lib.sce_workspace_bytes(C.byref(desc)) > 0
    desc.arith = 7
    assert lib.sce_workspace_bytes(C.byref(desc)) == 0
    desc.arith = _lib.SCE_ARITH_AUTO
    desc.d = 64
    desc.fwd_passes = 2
    assert lib.sce_workspace_bytes(C.byref(desc)) == 0
    plan = C.c_void_p()
    desc.fwd_passes = 3
    bufs = _lib.SceBuffers()                                         # all NULL
    assert lib.sce_plan_create(C.byref(desc), C.byref(bufs), C.byref(plan)) == -1
    assert b"required" in lib.sce_last_error()
    assert lib.sce_plan_destroy(None) == 0
    assert lib.sce_step(None, None, 1, None, None, None) == -1


def test_workspace_size_config2():
    """Config 2 (M=16, d=512, n=4096, B=8192): the (hi, lo) bf16 code + code gradient dominate: 4 x 1 GiB."""
    lib = _lib.load()
    desc = _lib.SceDesc(variant=0, n_models=16, d=512, n=4096, batch_max=8192, x_per_model=0, lr=1e-3, beta1=0.9,
                        beta2=0.999, eps=1e-8, eps_root=0.0, adam_count_mode=0, fwd_passes=3, bwd_passes=3,
                        norm_floor=1e-8)
    need = lib.sce_workspace_bytes(C.byref(desc))
    assert 4.5 * 2**30 < need < 5.5 * 2**30


def test_no_cpu_fallback():
    models = [S.FunctionalTiedSAE.init(16, 32, 1e-3) for _ in range(2)]
    ens = S.FunctionalEnsemble(models, S.FunctionalTiedSAE, S.adam, {"lr": 1e-3}, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        ens.step_batch(torch.randn(8, 16))
    with pytest.raises(RuntimeError, match="CUDA"):
        S.FunctionalTiedSAE.loss(*models[0], torch.randn(8, 16))

    class Custom(S.DictSignature):
        pass

    with pytest.raises(NotImplementedError, match="no engine variant"):
        S.FunctionalEnsemble(models, Custom, S.adam, {"lr": 1e-3}, device="cpu")


def test_init_matches_reference_initialisers():
    """Same RNG consumption as the reference's init (xavier encoder, zero bias, [decoder]) => identical tensors
    to the golden fixtures' parameters for the same seed (tests/golden/tied_small.pt was made with seed 0)."""
    fx = torch.load(os.path.join(ROOT, "tests", "golden", "tied_small.pt"), weights_only=False)
    torch.manual_seed(0)
    for i, l1 in enumerate([1e-3, 3e-3, 1e-2]):
        p, b = S.FunctionalTiedSAE.init(32, 64, l1, dtype=torch.float32)
        assert torch.equal(p["encoder"], fx["params"]["encoder"][i])
        assert torch.equal(p["encoder_bias"], fx["params"]["encoder_bias"][i])
        assert float(b["l1_alpha"]) == pytest.approx(l1)
        assert set(b) == {"center_rot", "center_trans", "center_scale", "l1_alpha", "bias_decay"}
    p, b = S.FunctionalSAE.init(8, 16, 1e-3, bias_decay=0.5)
    assert set(p) == {"encoder", "encoder_bias", "decoder"} and float(b["bias_decay"]) == 0.5
    p, b = S.FunctionalMaskedTiedSAE.init(8, 16, 32, 1e-3)
    assert p["encoder"].shape == (32, 8) and int(b["dict_size"]) == 16
    assert b["coef_mask"].tolist() == [False] * 16 + [True] * 16
    p, b = S.TopKEncoder.init(8, 16, 4)
    assert set(p) == {"dict"} and b["sparsity"].dtype == torch.long


def test_stack_unstack_state_dict_roundtrip():
    models = [S.FunctionalSAE.init(8, 16, a, bias_decay=0.1 * i) for i, a in enumerate((1e-3, 1e-2, 1e-1))]
    ens = S.FunctionalEnsemble(models, S.FunctionalSAE, "adam", {"lr": 3e-4}, device="cpu")
    assert ens.n_models == 3 and ens.params["encoder"].shape == (3, 16, 8)
    assert ens.buffers["l1_alpha"].shape == (3,)
    assert ens.optimizer.lr == pytest.approx(3e-4)
    back = ens.unstack(device="cpu")
    for (p0, b0), (p1, b1) in zip(models, back):
        assert all(torch.equal(p0[k], p1[k]) for k in p0) and all(torch.equal(b0[k], b1[k]) for k in b0)
    sd = ens.state_dict()
    for key in ("device", "n_models", "params", "buffers", "sig", "no_stacking", "optimizer_func", "optimizer_kwargs",
                "optim_states"):                                   # the reference's keys (ensemble.py:150-161)
        assert key in sd
    blob = io.BytesIO()
    torch.save(sd, blob)                                            # must survive pickling (mp spawn)
    blob.seek(0)
    ens2 = S.FunctionalEnsemble.from_state(torch.load(blob, weights_only=False))
    assert torch.equal(ens2.params["decoder"], ens.params["decoder"]) and ens2.sig is S.FunctionalSAE
    ens.to_shared_memory()
    assert ens.params["encoder"].is_shared()


def test_optimizer_resolution():
    from sparse_coding_b200.optim import resolve_optimizer
    assert resolve_optimizer("adam", {"lr": 1e-2}).lr == pytest.approx(1e-2)
    cfg = resolve_optimizer(S.adam, {"lr": 1e-3, "betas": (0.8, 0.9), "eps": 1e-6})
    assert (cfg.b1, cfg.b2, cfg.eps) == (pytest.approx(0.8), pytest.approx(0.9), pytest.approx(1e-6))
    with pytest.raises(ValueError):
        resolve_optimizer("sgd", {"lr": 1e-3})
    with pytest.raises(NotImplementedError):
        S.adam(lr=1e-3, weight_decay=0.1)
    with pytest.raises(ValueError):
        S.optim_str_to_func("sgd")
    assert S.optim_str_to_func("adam") is S.adam


def test_learned_dicts_match_reference(golden):
    fx = golden("learned_dicts")
    tied = S.TiedSAE(fx["encoder"], fx["bias"], centering=(fx["trans"], fx["rot"], fx["scale"]), norm_encoder=True)
    untied = S.UntiedSAE(fx["encoder"], fx["decoder"], fx["bias"])
    tk = S.TopKLearnedDict(fx["decoder"] / fx["decoder"].norm(dim=-1, keepdim=True), fx["topk_k"])
    X = fx["batch"]
    tol = dict(rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(tied.encode(tied.center(X)), fx["tied_encode"], **tol)
    torch.testing.assert_close(tied.predict(X), fx["tied_predict"], **tol)
    torch.testing.assert_close(tied.get_learned_dict(), fx["tied_dict"], **tol)
    torch.testing.assert_close(untied.encode(X), fx["untied_encode"], **tol)
    torch.testing.assert_close(untied.predict(X), fx["untied_predict"], **tol)
    torch.testing.assert_close(untied.get_learned_dict(), fx["untied_dict"], **tol)
    torch.testing.assert_close(tk.encode(X), fx["topk_encode"], **tol)
    torch.testing.assert_close(tk.predict(X), fx["topk_predict"], **tol)
    assert (tied.n_feats, tied.activation_size) == (64, 32) and tied.n_dict_components() == 64
    torch.testing.assert_close(tied.uncenter(tied.center(X)), X, rtol=1e-4, atol=1e-5)


def test_checkpoint_pickles_use_reference_names(tmp_path):
    """learned_dicts.pt written here must resolve inside the reference repo: classes pickle as
    autoencoders.learned_dict.* / autoencoders.topk_encoder.* (big_sweep.py:378-384 layout)."""
    from sparse_coding_b200.train_loop import unstacked_to_learned_dicts
    models = [S.FunctionalTiedSAE.init(8, 16, a) for a in (1e-3, 1e-2)]
    ens = S.FunctionalEnsemble(models, S.FunctionalTiedSAE, S.adam, {"lr": 1e-3}, device="cpu")
    dicts = unstacked_to_learned_dicts(ens, {"dict_size": 16, "device": "cpu"}, ["dict_size"], ["l1_alpha"])
    assert [h["dict_size"] for _, h in dicts] == [16, 16]
    assert dicts[1][1]["l1_alpha"] == pytest.approx(1e-2)
    path = tmp_path / "learned_dicts.pt"
    torch.save(dicts, path)
    raw = open(path, "rb").read()
    assert b"autoencoders.learned_dict" in raw and b"sparse_coding_b200" not in raw
    loaded = torch.load(path, weights_only=False)
    ld, hp = loaded[0]
    assert type(ld).__name__ == "TiedSAE" and ld.norm_encoder and ld.encoder.shape == (16, 8)
    import autoencoders.learned_dict as shim
    assert isinstance(ld, shim.TiedSAE)
    tk = S.TopKEncoder.to_learned_dict(*S.TopKEncoder.init(8, 16, 3))
    assert b"autoencoders.topk_encoder" in pickle.dumps(tk)
    with pytest.raises(ValueError, match="not found in args"):
        unstacked_to_learned_dicts(ens, {}, ["dict_size"], [])


def test_hyperparam_names():
    from sparse_coding_b200.train_loop import format_hyperparam_val, make_hyperparam_name
    assert format_hyperparam_val(1e-3) == "1.00E-03"
    assert format_hyperparam_val(2048.0) == "2.05E03"
    assert format_hyperparam_val(4096) == "4096"
    assert make_hyperparam_name({"dict_size": 4096, "l1_alpha": 3e-4}) == "dict_size_4096_l1_alpha_3.00E-04"


def test_batch_index_lists_follow_the_reference_sampler():
    """The one-shot permutation must be exactly what BatchSampler(RandomSampler) yields batch by batch."""
    from sparse_coding_b200.train_loop import _batch_index_lists
    N, B = 1000, 128
    mk = lambda: torch.utils.data.BatchSampler(torch.utils.data.RandomSampler(range(N)), batch_size=B, drop_last=False)
    torch.manual_seed(0)
    ref = [list(b) for b in mk()]
    torch.manual_seed(0)
    got = [t.tolist() for t in _batch_index_lists(mk())]
    assert got == ref and len(got[-1]) == N - (N // B) * B
    torch.manual_seed(0)
    again = [t.tolist() for t in _batch_index_lists(mk())]
    assert again == ref                                             # Q7: same shuffle after re-seeding


def test_resume_state_keeps_engine_settings(tmp_path):
    """save_resume_state / load_resume_state carry every engine-only setting: a run pinned to arith='bf16x3' (its
    activations leave the fp16 range) must not silently resume on the narrower f16f8 arithmetic, and a bf16x3
    fall-back taken by an 'auto' plan sticks as well."""
    from sparse_coding_b200.train_loop import load_resume_state, save_resume_state
    torch.manual_seed(0)
    models = [S.FunctionalTiedSAE.init(16, 32, a) for a in (1e-3, 1e-2)]
    ens = S.FunctionalEnsemble(models, S.FunctionalTiedSAE, S.adam, {"lr": 2e-3}, device="cpu", arith="bf16x3",
                               adam_count_mode="standard", bwd_passes=1, health_check_every=5)
    ens._steps = 7
    save_resume_state(ens, str(tmp_path / "a.pt"))
    back = load_resume_state(str(tmp_path / "a.pt"), "cpu")
    assert (back.arith, back.adam_count_mode, back.bwd_passes, back.fwd_passes) == ("bf16x3", "standard", 1, 3)
    assert back._steps == 7 and back.health_check_every == 5 and back.optimizer.lr == pytest.approx(2e-3)
    assert back._arith_fallback is None
    auto = S.FunctionalEnsemble(models, S.FunctionalTiedSAE, S.adam, {"lr": 1e-3}, device="cpu")
    auto._arith_fallback = "bf16x3"                     # what check_health() records after an fp16 overflow
    save_resume_state(auto, str(tmp_path / "b.pt"))
    back = load_resume_state(str(tmp_path / "b.pt"), "cpu")
    assert back.arith == "auto" and back._arith_fallback == "bf16x3"


def test_from_state_continues_the_step_count_of_a_dead_worker():
    """cluster_runs.py:113-125: the parent hands state_dict() to a freshly spawned worker per chunk; the worker's
    Python-side step counter dies with it, but optim_states['count'] is shared memory updated in place — with
    adam_count_mode='standard' the next worker continues the bias correction from there."""
    torch.manual_seed(0)
    models = [S.FunctionalTiedSAE.init(16, 32, 1e-3) for _ in range(2)]
    ens = S.FunctionalEnsemble(models, S.FunctionalTiedSAE, S.adam, {"lr": 1e-3}, device="cpu",
                               adam_count_mode="standard")
    sd = ens.state_dict()                               # the parent's view: steps == 0 for ever
    for t in ens.optim_states["count"].values():
        t.add_(12)                                      # what 12 steps of a child did to the shared tensors
    assert S.FunctionalEnsemble.from_state(sd)._steps == 12
    frozen = S.FunctionalEnsemble(models, S.FunctionalTiedSAE, S.adam, {"lr": 1e-3}, device="cpu")
    assert S.FunctionalEnsemble.from_state(frozen.state_dict())._steps == 0


def test_permutation_cache_changes_nothing():
    """_batch_index_lists with a cache returns exactly the batches it returns without one (the reference re-seeds the
    global RNG per chunk, so equal-length chunks draw the same permutation: SURVEY Q7), also after the cache was filled
    under another seed or chunk length."""
    from sparse_coding_b200.train_loop import _batch_index_lists
    mk = lambda n: torch.utils.data.BatchSampler(torch.utils.data.RandomSampler(range(n)), batch_size=64, drop_last=False)
    cache = {}
    for seed, n in ((0, 1000), (0, 1000), (5, 1000), (0, 777), (0, 1000)):
        torch.manual_seed(seed)
        plain = [b.clone() for b in _batch_index_lists(mk(n))]
        torch.manual_seed(seed)
        cached = [b.clone() for b in _batch_index_lists(mk(n), None, cache)]
        torch.manual_seed(seed)
        ref = [torch.tensor(ix) for ix in mk(n)]                      # what the reference's loop iterates over
        assert len(plain) == len(cached) == len(ref) == -(-n // 64)
        assert all(torch.equal(a, b) and torch.equal(a, c) for a, b, c in zip(plain, cached, ref))
    assert 1 <= len(cache) <= 4


def test_chunk_file_record_offset(tmp_path):
    """ChunkStreamer reads a chunk's bytes with read() straight into pinned memory when the file is a plain
    torch.save of one contiguous tensor (the reference's {i}.pt format): the data record's offset must point at
    exactly the tensor's bytes; anything else (several storages, a view of a larger storage) reports None -> mmap path."""
    from sparse_coding_b200.train_loop import _single_record_offset
    t = torch.randn(513, 40).half()
    p = str(tmp_path / "0.pt")
    torch.save(t, p)
    off = _single_record_offset(p, t.numel() * 2)
    raw = open(p, "rb").read()
    assert off is not None and torch.equal(torch.frombuffer(bytearray(raw[off:off + t.numel() * 2]), dtype=torch.float16).view(513, 40), t)
    torch.save([t, t + 1], p)
    assert _single_record_offset(p, t.numel() * 2) is None
    torch.save(t[:7], p)
    assert _single_record_offset(p, 7 * 40 * 2) is None
    assert _single_record_offset(str(tmp_path / "missing.pt"), 16) is None


def test_nvtx_ranges_are_off_by_default_and_switchable(monkeypatch):
    """SURVEY §5 tracing: SCE_NVTX=1 turns the ranges on at import; by default a range is a no-op context."""
    import contextlib
    import importlib
    import sparse_coding_b200.tracing as tr
    monkeypatch.delenv("SCE_NVTX", raising=False)
    tr = importlib.reload(tr)
    assert tr.ENABLED is False and isinstance(tr.nvtx_range("x"), contextlib.nullcontext)
    monkeypatch.setenv("SCE_NVTX", "1")
    tr = importlib.reload(tr)
    assert tr.ENABLED is True and not isinstance(tr.nvtx_range("x"), contextlib.nullcontext)
    monkeypatch.delenv("SCE_NVTX", raising=False)
    importlib.reload(tr)
