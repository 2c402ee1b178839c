"""Activation harvesting: chunks written by sparse_coding_b200.harvest must equal, bit for bit, what
the reference's hook-and-concatenate procedure (activation_dataset.py:441-503) produces on the same model/tokens,
in the same `{folder}/{tensor_name}/{i}.pt` fp16 layout. A tiny randomly initialised GPT-NeoX (the Pythia
architecture) stands in for the pretrained models, which cannot be downloaded here."""
import os

import pytest
import torch


def _tiny_lm():
    transformers = pytest.importorskip("transformers")
    cfg = transformers.GPTNeoXConfig(vocab_size=300, hidden_size=64, num_hidden_layers=3, num_attention_heads=4,
                                     intermediate_size=128, max_position_embeddings=32)
    torch.manual_seed(0)
    return transformers.GPTNeoXForCausalLM(cfg).eval()


def _reference_style(model, tokens, names, batch, dtype):
    """Restatement of the reference procedure: hook -> (b l) d -> cast -> cpu -> list, concatenated at the end."""
    bufs = {n: [] for n in names}
    handles = []
    mods = dict(model.named_modules())
    for n in names:
        handles.append(mods[n].register_forward_hook(
            lambda m, i, o, n=n: bufs[n].append((o[0] if isinstance(o, tuple) else o).reshape(-1, o[0].shape[-1] if isinstance(o, tuple) else o.shape[-1]).to(dtype).cpu())))
    with torch.no_grad():
        for i in range(0, tokens.shape[0], batch):
            model(tokens[i:i + batch].to(next(model.parameters()).device))
    for h in handles:
        h.remove()
    return {n: torch.cat(v) for n, v in bufs.items()}


def _run(device, tmp_path):
    from sparse_coding_b200.harvest import make_activation_dataset_hf
    model = _tiny_lm().to(device)
    L, bs = 16, 4
    tokens = torch.randint(0, 300, (40, L), generator=torch.Generator().manual_seed(1))
    dataset = [{"input_ids": t} for t in tokens]
    names = ["gpt_neox.layers.1", "gpt_neox.layers.2.mlp"]
    out = str(tmp_path / "acts")
    written = make_activation_dataset_hf(dataset, model, names, chunk_size=3 * bs * L + 5, n_chunks=10, output_folder=out,
                                         device=torch.device(device), max_length=L, model_batch_size=bs,
                                         precision="float16")
    ref = _reference_style(model, tokens, names, bs, torch.float16)
    rows_per_chunk = 3 * bs * L
    for n in names:
        files = sorted(os.listdir(os.path.join(out, n)), key=lambda f: int(f[:-3]))
        assert files == ["0.pt", "1.pt", "2.pt", "3.pt"] and len(written[n]) == 4   # 10 batches: 3 + 3 + 3 + 1
        got = torch.cat([torch.load(os.path.join(out, n, f)) for f in files])
        assert got.dtype == torch.float16 and got.shape == (40 * L, 64)
        assert torch.equal(got, ref[n])
        assert torch.load(os.path.join(out, n, "0.pt")).shape[0] == rows_per_chunk
        assert torch.load(os.path.join(out, n, "3.pt")).shape[0] == bs * L                # undersized final chunk
    return out, names


def test_harvest_matches_reference_procedure_cpu(tmp_path):
    _run("cpu", tmp_path)


def test_harvest_argument_errors(tmp_path):
    from sparse_coding_b200.harvest import make_activation_dataset_hf
    model = _tiny_lm()
    data = [{"input_ids": torch.zeros(8, dtype=torch.long)}]
    with pytest.raises(ValueError, match="precision"):
        make_activation_dataset_hf(data, model, ["gpt_neox.layers.0"], 64, 1, str(tmp_path), device="cpu", max_length=8,
                                   model_batch_size=1, precision="bfloat16")
    with pytest.raises(KeyError):
        make_activation_dataset_hf(data, model, ["nope"], 64, 1, str(tmp_path), device="cpu", max_length=8,
                                   model_batch_size=1)
    with pytest.raises(ValueError, match="smaller"):
        make_activation_dataset_hf(data, model, ["gpt_neox.layers.0"], 4, 1, str(tmp_path), device="cpu", max_length=8,
                                   model_batch_size=1)


@pytest.mark.gpu
def test_harvest_on_gpu_then_train(tmp_path):
    """GPU harvest (device chunk buffers, async D2H) == reference procedure; the chunks then feed the engine."""
    import sparse_coding_b200 as S
    from sparse_coding_b200.train_loop import train_on_chunks
    out, names = _run("cuda", tmp_path)
    folder = os.path.join(out, names[0])
    torch.manual_seed(0)
    models = [S.FunctionalTiedSAE.init(64, 128, a) for a in (1e-3, 1e-2)]
    ens = S.FunctionalEnsemble(models, S.FunctionalTiedSAE, S.adam, {"lr": 1e-3}, device="cuda")
    dicts = train_on_chunks(ens, {"device": "cuda", "dict_size": 128}, folder, str(tmp_path / "sweep"), 64,
                            ["dict_size"], ["l1_alpha"], chunk_order=[0, 1, 2, 3])
    assert len(dicts) == 2 and all(torch.isfinite(ld.encoder).all() for ld, _ in dicts)


# ----------------------------------------------------------------------------------------------------------------------
# Golden fixtures written by the REFERENCE's own make_activation_dataset_tl / make_activation_dataset
# (oracle/make_harvest_golden.py): chunk boundaries, flattening, first-chunk centring, the baukit GELU.
# The language model's forward pass is recomputed here, so values are compared to one fp16 ulp (a different CPU or
# the GPU may round an fp32 activation to the neighbouring fp16 value); shapes and chunk counts must match exactly.
# ----------------------------------------------------------------------------------------------------------------------
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "harvest.pt")
TL_CASES = ["tl_resid", "tl_resid_centred", "tl_mlp_skip1", "tl_attn_concat"]
SINGLE_CASES = ["single_tl", "single_tl_centred", "single_baukit", "single_baukit_centred", "single_tl_2chunks"]


def _fixture(device):
    pytest.importorskip("transformers")
    from oracle import harvest_models as HM
    fx = torch.load(GOLDEN)
    lm = HM.tiny_neox()
    lm.load_state_dict(fx["lm_state"])
    nano = HM.TinyNano()
    nano.load_state_dict(fx["nano_state"])
    rows = [{"input_ids": t} for t in fx["tokens"]]
    loader = torch.utils.data.DataLoader(rows, batch_size=fx["model_batch_size"], shuffle=False)
    return fx, HM.TinyHooked(lm).to(device), nano.to(device), loader


def _read(folder):
    files = sorted(os.listdir(folder), key=lambda f: int(f[:-3]))
    assert files == [f"{i}.pt" for i in range(len(files))]
    return [torch.load(os.path.join(folder, f)) for f in files]


def _same_chunks(got, want):
    assert [tuple(t.shape) for t in got] == [tuple(t.shape) for t in want]
    for g, w in zip(got, want):
        assert g.dtype == w.dtype == torch.float16 and g.is_contiguous()
        assert g.untyped_storage().nbytes() == g.numel() * 2          # compact file, also for undersized chunks
        torch.testing.assert_close(g.float(), w.float(), rtol=2e-3, atol=2e-4)


def _tl_case(tag, device, tmp_path):
    from sparse_coding_b200.harvest import make_activation_dataset_tl
    fx, hooked, _, loader = _fixture(device)
    c = fx["cases"][tag]
    folders = [str(tmp_path / tag / str(l)) for l in c["layers"]]
    n_act = make_activation_dataset_tl(loader, hooked, c["activation_width"], folders, layers=c["layers"],
                                       tensor_loc=c["tensor_loc"], chunk_size_gb=c["chunk_size_gb"],
                                       device=torch.device(device), n_chunks=c["n_chunks"], max_length=fx["max_length"],
                                       model_batch_size=fx["model_batch_size"], skip_chunks=c["skip_chunks"],
                                       center_dataset=c["center_dataset"])
    assert n_act == c["n_activations"]
    for folder, want in zip(folders, c["chunks"]):
        _same_chunks(_read(folder), want)


def _single_case(tag, device, tmp_path):
    from sparse_coding_b200.harvest import make_activation_dataset
    fx, hooked, nano, loader = _fixture(device)
    c = fx["cases"][tag]
    folder = str(tmp_path / tag)
    make_activation_dataset(loader, nano if c["baukit"] else hooked, c["tensor_name"], c["activation_width"], folder,
                            baukit=c["baukit"], chunk_size_gb=c["chunk_size_gb"], device=torch.device(device),
                            layer=c["layer"], n_chunks=c["n_chunks"], max_length=fx["max_length"],
                            model_batch_size=fx["model_batch_size"], center_dataset=c["center_dataset"])
    _same_chunks(_read(folder), c["chunks"])


@pytest.mark.parametrize("tag", TL_CASES)
def test_tl_harvest_matches_reference_files_cpu(tag, tmp_path):
    _tl_case(tag, "cpu", tmp_path)


@pytest.mark.parametrize("tag", SINGLE_CASES)
def test_single_tensor_harvest_matches_reference_files_cpu(tag, tmp_path):
    _single_case(tag, "cpu", tmp_path)


def test_tensor_names_match_reference():
    from sparse_coding_b200.harvest import make_tensor_name
    from oracle import harvest_models as HM
    fx = torch.load(GOLDEN)
    for loc in ("residual", "mlp", "attn", "attn_concat", "mlpout"):
        assert make_tensor_name(3, loc, HM.TINY_TL_NAME) == fx["tensor_names"][loc]
    assert make_tensor_name(3, "mlp", "nanoGPT") == fx["tensor_names"]["mlp_nanoGPT"]
    with pytest.raises(NotImplementedError):
        make_tensor_name(3, "residual", "nanoGPT")
    with pytest.raises(AssertionError):
        make_tensor_name(3, "nowhere", HM.TINY_TL_NAME)


def test_tl_harvest_stops_cleanly_on_a_chunk_boundary(tmp_path):
    """8 model batches in chunks of 3 + 1: the reference goes on to torch.cat([]) (activation_dataset.py:382, :500);
    here the run ends after the two full chunks."""
    from sparse_coding_b200.harvest import make_activation_dataset_tl
    fx, hooked, _, _ = _fixture("cpu")
    rows = [{"input_ids": t} for t in fx["tokens"][:32]]
    loader = torch.utils.data.DataLoader(rows, batch_size=4, shuffle=False)
    c = fx["cases"]["tl_resid"]
    folder = str(tmp_path / "b")
    n = make_activation_dataset_tl(loader, hooked, c["activation_width"], [folder], layers=[1],
                                   chunk_size_gb=c["chunk_size_gb"], device=torch.device("cpu"), n_chunks=5,
                                   max_length=fx["max_length"], model_batch_size=4)
    assert n == 32 * fx["max_length"] and [t.shape[0] for t in _read(folder)] == [256, 256]


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["tl_resid_centred", "tl_attn_concat"])
def test_tl_harvest_matches_reference_files_gpu(tag, tmp_path):
    _tl_case(tag, "cuda", tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["single_tl_centred", "single_baukit_centred"])
def test_single_tensor_harvest_matches_reference_files_gpu(tag, tmp_path):
    _single_case(tag, "cuda", tmp_path)
