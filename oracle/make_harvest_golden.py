"""Generate tests/golden/harvest.pt by running the REFERENCE's own harvesting functions
(HoagyC/sparse_coding @ 69c5ae0, activation_dataset.py): ``make_activation_dataset_tl`` (:323-391), ``make_activation_dataset`` (:263-321, both its TransformerLens and its
baukit branch) and ``save_activation_chunk`` (:499-503), on the tiny models of oracle/harvest_models.py.

TEST INFRASTRUCTURE. Run in the build container only (needs /root/reference):
    python oracle/make_harvest_golden.py

``activation_dataset`` imports packages that are not installed here. They are replaced by stubs that carry no
harvesting logic:
* ``boto3`` / ``botocore`` (pulled in by ``from utils import *``): inert.
* ``transformer_lens``: ``HookedTransformer`` is only a type annotation in the functions run here;
  ``get_official_model_name`` knows exactly one name (the tiny model's) and raises ``ValueError`` otherwise, which is
  all ``check_transformerlens_model`` (:61-66) asks of it.
* ``baukit.Trace``: the published behaviour the reference relies on (:291-293) — a context manager that records the
  output of the named submodule during the forward pass as ``.output``.
Everything recorded below — the cast, the flattening, the chunk boundaries, the centring, the files — is produced by
the reference's code. The fixture stores the models' weights, the token rows, and every chunk file's tensor.
"""
import os
import sys
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
REF = os.environ.get("SCE_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "..", "tests", "golden", "harvest.pt")

from oracle import harvest_models as HM  # noqa: E402


def import_reference():
    for name in ("boto3", "botocore"):
        sys.modules.setdefault(name, types.ModuleType(name))
    exc = types.ModuleType("botocore.exceptions")
    exc.ClientError = type("ClientError", (Exception,), {})
    exc.NoCredentialsError = type("NoCredentialsError", (Exception,), {})
    sys.modules.setdefault("botocore.exceptions", exc)

    tl = types.ModuleType("transformer_lens")
    tl.HookedTransformer = type("HookedTransformer", (), {})
    lfp = types.ModuleType("transformer_lens.loading_from_pretrained")

    def get_official_model_name(name):
        if name == HM.TINY_TL_NAME:
            return name
        raise ValueError(f"{name} not found")

    lfp.get_official_model_name = get_official_model_name
    lfp.convert_hf_model_config = lambda name: dict(d_model=HM.D_MODEL, d_mlp=HM.D_MLP, n_heads=HM.N_HEADS,
                                                    d_head=HM.D_MODEL // HM.N_HEADS)
    tl.loading_from_pretrained = lfp
    sys.modules.setdefault("transformer_lens", tl)
    sys.modules.setdefault("transformer_lens.loading_from_pretrained", lfp)

    bk = types.ModuleType("baukit")

    class Trace:
        def __init__(self, module, layer):
            self.output = None
            target = dict(module.named_modules())[layer]
            self._h = target.register_forward_hook(lambda m, i, o: setattr(self, "output", o))

        def __enter__(self):
            return self

        def __exit__(self, *a):
            self._h.remove()

    bk.Trace = Trace
    sys.modules.setdefault("baukit", bk)
    sys.path.insert(0, REF)
    import activation_dataset as AD
    return AD


def read_chunks(folder):
    files = sorted(os.listdir(folder), key=lambda f: int(f[:-3]))
    return [torch.load(os.path.join(folder, f)) for f in files]


def main():
    AD = import_reference()
    cpu = torch.device("cpu")
    L, bs, n_sent = HM.CTX, 4, 38          # 38 sentences: 9 full model batches + one of 2 rows
    tokens = torch.randint(0, HM.VOCAB, (n_sent, L), generator=torch.Generator().manual_seed(1))
    rows = [{"input_ids": t} for t in tokens]
    loader = lambda: torch.utils.data.DataLoader(rows, batch_size=bs, shuffle=False)   # noqa: E731
    lm = HM.tiny_neox()
    hooked = HM.TinyHooked(lm)
    nano = HM.TinyNano()
    fx = dict(tokens=tokens, model_batch_size=bs, max_length=L, lm_state=lm.state_dict(), nano_state=nano.state_dict(),
              cases={})

    def gb(width, batches):       # chunk_size_gb that makes `chunk_size // activation_size` equal `batches`
        return (batches + 0.5) * (width * 2 * bs * L) / 2 ** 30

    with tempfile.TemporaryDirectory() as tmp:
        # --- HF hooks (:393-497): NOT recordable. The reference registers `hook(module, output, tensor_name=...)` with
        # `register_forward_hook` (:443-454), which calls hooks as (module, input, output): the module's output arrives
        # in `tensor_name` and the lookup `tensor_buffer[tensor_name]` raises KeyError on the first forward pass, with
        # any torch version. (Its chunk-boundary test `batch_idx+1 % chunk_batches == 0`, :466, would never fire
        # either.) sparse_coding_b200.harvest.make_activation_dataset_hf implements the evident intent and is tested
        # against a restatement of it (tests/test_harvest.py); the fixtures below pin the variants that do run.
        # --- TransformerLens cache, several layers at once (:323-391); chunks hold max_batches_per_chunk + 1 batches
        # (`batch_idx >= max_batches_per_chunk`, :374)
        # (the reference only terminates cleanly when the data runs out inside a chunk that then holds FEWER than
        # max_batches_per_chunk batches, or when n_chunks is reached: otherwise it goes on to torch.cat([]) (:382, :500).
        # 10 batches in chunks of 3 + 1 end as 4 + 4 + 2; the skip_chunks case stops at n_chunks = 1.)
        for tag, loc, width, centre, skip, n_chunks in (("tl_resid", "residual", HM.D_MODEL, False, 0, 5),
                                                        ("tl_resid_centred", "residual", HM.D_MODEL, True, 0, 5),
                                                        ("tl_mlp_skip1", "mlp", HM.D_MLP, False, 1, 1),
                                                        ("tl_attn_concat", "attn_concat", HM.D_MODEL, True, 0, 2)):
            layers = [0, 2]
            folders = [os.path.join(tmp, tag, str(l)) for l in layers]
            n_act = AD.make_activation_dataset_tl(loader(), hooked, width, folders, layers=layers, tensor_loc=loc,
                                                  chunk_size_gb=gb(width, 3), device=cpu, n_chunks=n_chunks, max_length=L,
                                                  model_batch_size=bs, skip_chunks=skip, center_dataset=centre)
            fx["cases"][tag] = dict(layers=layers, tensor_loc=loc, activation_width=width, chunk_size_gb=gb(width, 3),
                                    n_chunks=n_chunks, skip_chunks=skip, center_dataset=centre, n_activations=n_act,
                                    chunks=[read_chunks(f) for f in folders])

        # --- single tensor, TransformerLens branch and baukit branch (:263-321)
        for tag, model, baukit, name, width, layer, centre in (
                ("single_tl", hooked, False, "blocks.1.hook_resid_post", HM.D_MODEL, 1, False),
                ("single_tl_centred", hooked, False, "blocks.1.hook_mlp_out", HM.D_MODEL, 1, True),
                ("single_baukit", nano, True, "transformer.h.1.mlp.c_fc", HM.D_MLP, 1, False),
                ("single_baukit_centred", nano, True, "transformer.h.2.mlp.c_fc", HM.D_MLP, 2, True)):
            folder = os.path.join(tmp, tag)
            AD.make_activation_dataset(loader(), model, name, width, folder, baukit=baukit, chunk_size_gb=gb(width, 4),
                                       device=cpu, layer=layer, n_chunks=5, max_length=L, model_batch_size=bs,
                                       center_dataset=centre)
            fx["cases"][tag] = dict(tensor_name=name, activation_width=width, baukit=baukit, layer=layer,
                                    chunk_size_gb=gb(width, 4), n_chunks=5, center_dataset=centre,
                                    chunks=read_chunks(folder))
        # n_chunks reached before the data runs out
        folder = os.path.join(tmp, "single_tl_2chunks")
        AD.make_activation_dataset(loader(), hooked, "blocks.0.hook_resid_post", HM.D_MODEL, folder, baukit=False,
                                   chunk_size_gb=gb(HM.D_MODEL, 4), device=cpu, layer=0, n_chunks=2, max_length=L,
                                   model_batch_size=bs)
        fx["cases"]["single_tl_2chunks"] = dict(tensor_name="blocks.0.hook_resid_post", activation_width=HM.D_MODEL,
                                                baukit=False, layer=0, chunk_size_gb=gb(HM.D_MODEL, 4), n_chunks=2,
                                                center_dataset=False, chunks=read_chunks(folder))

        fx["tensor_names"] = {loc: AD.make_tensor_name(3, loc, HM.TINY_TL_NAME)
                              for loc in ("residual", "mlp", "attn", "attn_concat", "mlpout")}
        fx["tensor_names"]["mlp_nanoGPT"] = AD.make_tensor_name(3, "mlp", "nanoGPT")

    torch.save(fx, OUT)
    for k, v in fx["cases"].items():
        ch = v["chunks"]
        flat = ch if isinstance(ch, list) and ch and torch.is_tensor(ch[0]) else (
            [t for c in (ch.values() if isinstance(ch, dict) else ch) for t in c])
        print(k, [tuple(t.shape) for t in flat], flat[0].dtype)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
