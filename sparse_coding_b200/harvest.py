"""Activation harvesting — the producers of the `{folder}/{i}.pt` fp16 chunks the sweep trains on (SURVEY §8 f4;
reference: ``activation_dataset.py``).

Three capture mechanisms, same names / arguments / file layout as the reference's:

* ``make_activation_dataset_tl``  (:323-391) — several layers at once out of a TransformerLens-style
  ``model.run_with_cache(tokens, stop_at_layer=...)``; what ``setup_data`` (:544-604) and so ``big_sweep`` use.
* ``make_activation_dataset``     (:263-321) — one tensor, either from ``run_with_cache`` or (``baukit=True``, the
  nanoGPT path) from a forward hook on the named module followed by the reference's GELU.
* ``make_activation_dataset_hf``  (:393-497) — forward hooks on named modules of a HF ``AutoModelForCausalLM``.

The reference casts every captured tensor to fp16, keeps a Python list of the pieces (moved to the host inside the
forward pass in the HF variant: one synchronous D2H copy per hooked module per model batch) and concatenates them at
every chunk boundary. Here a capture only copies (cast fused into the copy) into a preallocated DEVICE buffer of one
chunk; centring runs on that buffer; a full chunk leaves the device in a single asynchronous copy to pinned memory on
a side stream while the language model keeps running into the second buffer, and a background thread writes the
file. The on-disk format is unchanged: ``torch.save`` of one ``[rows, d]`` fp16 (or fp32) tensor per chunk. The
language model itself is whatever the caller passes; its forward pass is library code, not part of this engine.
TransformerLens and baukit are not dependencies: the model is duck-typed (``run_with_cache`` / named modules).

Parity: tests/golden/harvest.pt holds the chunk files the reference's own ``make_activation_dataset_tl`` and
``make_activation_dataset`` wrote for tiny models (oracle/make_harvest_golden.py), including their chunk-boundary
rules (a TransformerLens chunk holds ``max_batches_per_chunk + 1`` model batches, :374) and first-chunk centring.

Deliberate differences: (1) the reference's HF variant cannot run — its hook is registered with the wrong arity
(:443-454: the module output lands in ``tensor_name`` and the buffer lookup raises ``KeyError``), and its
chunk-boundary test ``batch_idx+1 % chunk_batches == 0`` (:466) never fires; the HF variant here records the module
output and cuts chunks every ``chunk_size // (model_batch_size * max_length)`` batches, as evidently intended.
(2) Where the reference runs into ``torch.cat([])`` because the data ended exactly on a chunk boundary (:382 / :318
with :500), the functions here stop without writing an empty chunk.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, Iterable, List, Optional

import torch


def save_activation_chunk(dataset: torch.Tensor, n_saved_chunks: int, dataset_folder: str) -> str:
    """One chunk file, reference layout (activation_dataset.py:499-503)."""
    os.makedirs(dataset_folder, exist_ok=True)
    path = os.path.join(dataset_folder, f"{n_saved_chunks}.pt")
    with open(path, "wb") as f:
        torch.save(dataset, f)
    return path


class _ChunkSink:
    """Two device buffers of one chunk each for one captured tensor; rows are appended as the model runs."""

    def __init__(self, rows: int, dtype: torch.dtype, device: torch.device, folder: str):
        self.rows, self.dtype, self.device, self.folder = rows, dtype, device, folder
        self.bufs: List[Optional[torch.Tensor]] = [None, None]
        self.pinned: List[Optional[torch.Tensor]] = [None, None]
        self.copied = [None, None]          # event: D2H out of buffer `slot` finished
        self.writer = [None, None]          # future: the file write that reads pinned buffer `slot`
        self.slot, self.fill, self.width = 0, 0, None
        self.pieces = 0                     # captures appended to the current chunk

    def append(self, out: torch.Tensor) -> torch.Tensor:
        """`b s n -> (b s) n` / `b s h d -> (b s) (h d)` (activation_dataset.py:304, :366-368), cast fused into the
        copy. Returns the rows just written (a view of the device buffer)."""
        flat = out.reshape(-1, out.shape[-1]) if out.dim() <= 3 else out.reshape(out.shape[0] * out.shape[1], -1)
        if self.width is None:
            self.width = flat.shape[1]
        for s in range(2):
            if self.bufs[s] is None:
                self.bufs[s] = torch.empty(self.rows, self.width, dtype=self.dtype, device=self.device)
        take = min(flat.shape[0], self.rows - self.fill)
        if self.fill == 0 and self.copied[self.slot] is not None and self.device.type == "cuda":
            # this buffer is being reused: its previous contents must have left for the host first
            torch.cuda.current_stream(self.device).wait_event(self.copied[self.slot])
        dst = self.bufs[self.slot][self.fill:self.fill + take]
        dst.copy_(flat[:take])
        self.fill += take
        self.pieces += 1
        # rows beyond the chunk boundary are dropped, as a chunk is a whole number of model batches
        return dst

    def rows_view(self) -> torch.Tensor:
        return self.bufs[self.slot][:self.fill]

    def full(self) -> bool:
        return self.fill >= self.rows


class _Shipper:
    """Moves finished chunks off the device: async D2H into pinned memory on a side stream, file write on a worker
    thread, while the model fills the sink's other buffer."""

    def __init__(self, device: torch.device):
        self.device = device
        self.on_gpu = device.type == "cuda"
        self.copy_stream = torch.cuda.Stream(device) if self.on_gpu else None
        self.pool = ThreadPoolExecutor(max_workers=1)
        self.pending: List = []             # (key, future) in submission order

    def flush(self, key, sink: _ChunkSink, chunk_idx: int) -> None:
        rows, slot = sink.fill, sink.slot
        src = sink.bufs[slot][:rows]
        if self.on_gpu:
            if sink.pinned[slot] is None:
                sink.pinned[slot] = torch.empty(sink.rows, sink.width, dtype=sink.dtype).pin_memory()
            if sink.writer[slot] is not None:
                sink.writer[slot].result()                              # the previous file out of this pinned buffer is written
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))   # the captures' copies are done
            with torch.cuda.stream(self.copy_stream):
                sink.pinned[slot][:rows].copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            sink.copied[slot] = ev
            host = sink.pinned[slot][:rows]

            def job(ev=ev, host=host, full=rows == sink.rows):
                ev.synchronize()
                # torch.save serialises the whole underlying storage of a view: an undersized chunk is written from a
                # compact copy (the reference writes a compact torch.cat result, activation_dataset.py:499-503)
                return save_activation_chunk(host if full else host.clone(), chunk_idx, sink.folder)
        else:
            host = src.clone()

            def job(host=host):
                return save_activation_chunk(host, chunk_idx, sink.folder)
        fut = self.pool.submit(job)
        sink.writer[slot] = fut
        self.pending.append((key, fut))
        sink.slot ^= 1
        sink.fill = 0
        sink.pieces = 0

    def finish(self) -> Dict:
        written: Dict = {}
        try:
            for key, fut in self.pending:
                written.setdefault(key, []).append(fut.result())
        finally:
            self.pool.shutdown()
        return written


def _precision(precision: str) -> torch.dtype:
    if precision == "float16":
        return torch.float16
    if precision == "float32":
        return torch.float32
    raise ValueError(f"Invalid precision '{precision}'")


def make_activation_dataset_hf(sentence_dataset, model: torch.nn.Module, tensor_names: List[str], chunk_size: int,
                               n_chunks: int, output_folder: str = "activation_data", skip_chunks: int = 0,
                               device: Optional[torch.device] = torch.device("cuda:0"), max_length: int = 2048,
                               model_batch_size: int = 4, precision: str = "float16",
                               shuffle_seed: Optional[int] = None) -> Dict[str, List[str]]:
    """Run ``model`` over ``sentence_dataset`` (items with an ``"input_ids"`` tensor of ``max_length`` tokens) and
    write, for every module name in ``tensor_names``, chunks of ``chunk_size`` activation rows to
    ``{output_folder}/{tensor_name}/{i}.pt``. Same signature and file layout as the reference
    (activation_dataset.py:393-405). Returns the written paths per tensor name."""
    dtype = _precision(precision)
    device = torch.device(device)
    chunk_batches = chunk_size // (model_batch_size * max_length)
    if chunk_batches < 1:
        raise ValueError("chunk_size is smaller than one model batch")
    rows_per_chunk = chunk_batches * model_batch_size * max_length
    if shuffle_seed is not None:
        torch.manual_seed(shuffle_seed)
    loader = torch.utils.data.DataLoader(sentence_dataset, batch_size=model_batch_size, shuffle=shuffle_seed is not None)
    it = iter(loader)
    for _ in range(skip_chunks * chunk_batches):
        next(it)

    modules = dict(model.named_modules())
    for name in tensor_names:
        if name not in modules:
            raise KeyError(f"module '{name}' not found in the model")
    sinks = {name: _ChunkSink(rows_per_chunk, dtype, device, os.path.join(output_folder, name)) for name in tensor_names}
    ship = _Shipper(device)
    handles = []
    for name in tensor_names:
        def hook(module, inputs, output, name=name):
            out = output[0] if isinstance(output, tuple) else output
            sinks[name].append(out.detach())
            return output

        handles.append(modules[name].register_forward_hook(hook))

    chunk_idx = 0
    batches_in_chunk = 0
    try:
        with torch.no_grad():
            model.eval()
            for batch in it:
                ids = batch["input_ids"].to(device)
                model(ids)
                batches_in_chunk += 1
                if batches_in_chunk == chunk_batches:
                    for name in tensor_names:
                        ship.flush(name, sinks[name], chunk_idx)
                    chunk_idx += 1
                    batches_in_chunk = 0
                    if chunk_idx >= n_chunks:
                        break
            if chunk_idx < n_chunks and batches_in_chunk > 0:           # undersized final chunk
                for name in tensor_names:
                    ship.flush(name, sinks[name], chunk_idx)
    finally:
        for h in handles:
            h.remove()
        written = ship.finish()
    return {name: written.get(name, []) for name in tensor_names}


# ------------------------------------------------------------------------------------------------------------------
# TransformerLens / baukit variants (the ones `setup_data`, activation_dataset.py:544-604, dispatches to)
# ------------------------------------------------------------------------------------------------------------------
_LAYER_LOCS = ("residual", "mlp", "attn", "attn_concat", "mlpout")


def make_tensor_name(layer: int, layer_loc: str, model_name: str) -> str:
    """Hook-point name of a layer location (activation_dataset.py:69-106). The reference asks TransformerLens'
    model registry whether ``model_name`` is one of its models; without that package every name other than
    ``"nanoGPT"`` is taken to be one (nanoGPT supports ``"mlp"`` only, as in the reference)."""
    assert layer_loc in _LAYER_LOCS, f"Layer location {layer_loc} not supported"
    if model_name == "nanoGPT":
        if layer_loc == "mlp":
            return f"transformer.h.{layer}.mlp.c_fc"
        raise NotImplementedError(f"Model {model_name} not supported for {layer_loc}")
    return {"residual": f"blocks.{layer}.hook_resid_post", "attn_concat": f"blocks.{layer}.attn.hook_z",
            "mlp": f"blocks.{layer}.mlp.hook_post", "attn": f"blocks.{layer}.hook_resid_post",
            "mlpout": f"blocks.{layer}.hook_mlp_out"}[layer_loc]


def _centre(sink: _ChunkSink, mean: Optional[torch.Tensor]) -> torch.Tensor:
    """Subtract the FIRST chunk's mean from the chunk in the sink's device buffer (activation_dataset.py:307-310,
    :378-381: ``torch.mean(torch.cat(dataset), dim=0)`` in the storage dtype, then ``x - chunk_mean``)."""
    rows = sink.rows_view()
    if mean is None:
        mean = torch.mean(rows, dim=0)
    rows.sub_(mean)
    return mean


def make_activation_dataset_tl(sentence_dataset: Iterable, model, activation_width: int, dataset_folders: List[str],
                               layers: List[int] = [2], tensor_loc: str = "residual", chunk_size_gb: float = 2,
                               device: torch.device = torch.device("cuda:0"), n_chunks: int = 1, max_length: int = 256,
                               model_batch_size: int = 4, skip_chunks: int = 0, center_dataset: bool = False) -> int:
    """Several layers of one model in one pass (activation_dataset.py:323-391): ``sentence_dataset`` is a DataLoader
    of ``{"input_ids": [model_batch_size, max_length]}`` batches, ``model`` anything with TransformerLens'
    ``run_with_cache(tokens, stop_at_layer=...) -> (logits, cache)`` and ``cfg.model_name``. Layer ``layers[i]``
    goes to ``dataset_folders[i]/{chunk}.pt`` as fp16 rows. A chunk holds ``max_batches_per_chunk + 1`` model batches
    (``batch_idx >= max_batches_per_chunk`` breaks AFTER the append, :374); the run ends after ``n_chunks`` chunks
    or at the first chunk with fewer than ``max_batches_per_chunk`` batches. Returns the number of activation rows
    seen for the first layer."""
    device = torch.device(device)
    chunk_size = chunk_size_gb * (2 ** 30)
    activation_size = activation_width * 2 * model_batch_size * max_length
    max_batches_per_chunk = int(chunk_size // activation_size)
    rows_cap = (max_batches_per_chunk + 1) * model_batch_size * max_length
    names = {layer: make_tensor_name(layer, tensor_loc, model.cfg.model_name) for layer in layers}
    sinks = {layer: _ChunkSink(rows_cap, torch.float16, device, folder) for layer, folder in zip(layers, dataset_folders)}
    ship = _Shipper(device)
    means: Dict[int, torch.Tensor] = {}
    it = iter(sentence_dataset)
    n_activations = 0
    try:
        with torch.no_grad():
            for _ in range(skip_chunks * max_batches_per_chunk):
                next(it)
            for chunk_idx in range(n_chunks):
                n_batches = 0
                for batch_idx, batch in enumerate(it):
                    tokens = batch["input_ids"].to(device)
                    _, cache = model.run_with_cache(tokens, stop_at_layer=max(layers) + 1)
                    for layer in layers:
                        piece = sinks[layer].append(cache[names[layer]])
                        if layer == layers[0]:
                            n_activations += piece.shape[0]
                    n_batches += 1
                    if batch_idx >= max_batches_per_chunk:
                        break
                if n_batches == 0:
                    break                                   # data ended on a chunk boundary (the reference: torch.cat([]))
                for layer in sinks:
                    if center_dataset:
                        means[layer] = _centre(sinks[layer], means.get(layer))
                    ship.flush(layer, sinks[layer], chunk_idx)
                if n_batches < max_batches_per_chunk:
                    break                                   # undersized chunk: the data ran out
    finally:
        ship.finish()
    return n_activations


def make_activation_dataset(sentence_dataset: Iterable, model, tensor_name: str, activation_width: int,
                            dataset_folder: str, baukit: bool = False, chunk_size_gb: float = 2,
                            device: torch.device = torch.device("cuda:0"), layer: int = 2, n_chunks: int = 1,
                            max_length: int = 256, model_batch_size: int = 4, center_dataset: bool = False) -> None:
    """One tensor of one model (activation_dataset.py:263-321). ``baukit=False``: ``cache[tensor_name]`` of
    ``model.run_with_cache(tokens, stop_at_layer=layer + 1)``. ``baukit=True`` (the nanoGPT path, :289-296): the
    output of the submodule called ``tensor_name`` during ``model(tokens)``, cast to fp16 and passed through GELU in
    fp16. Chunks of ``chunk_size // activation_size`` model batches; whatever is left when the data ends is written
    as a last, undersized chunk."""
    device = torch.device(device)
    chunk_size = chunk_size_gb * (2 ** 30)
    activation_size = activation_width * 2 * model_batch_size * max_length
    actives_per_chunk = chunk_size // activation_size
    if actives_per_chunk < 1:
        actives_per_chunk = 1                                # the reference saves after every batch in that case
    sink = _ChunkSink(int(actives_per_chunk) * model_batch_size * max_length, torch.float16, device, dataset_folder)
    ship = _Shipper(device)
    captured: List[torch.Tensor] = []
    handle = None
    if baukit:
        modules = dict(model.named_modules())
        if tensor_name not in modules:
            raise KeyError(f"module '{tensor_name}' not found in the model")
        handle = modules[tensor_name].register_forward_hook(
            lambda m, i, o: captured.append((o[0] if isinstance(o, tuple) else o).detach()))
    mean = None
    n_saved_chunks = 0
    try:
        with torch.no_grad():
            for batch in sentence_dataset:
                tokens = batch["input_ids"].to(device)
                if baukit:
                    captured.clear()
                    model(tokens)
                    piece = sink.append(captured[-1])
                    piece.copy_(torch.nn.functional.gelu(piece))
                else:
                    _, cache = model.run_with_cache(tokens, stop_at_layer=layer + 1)
                    sink.append(cache[tensor_name])
                if sink.pieces >= actives_per_chunk:
                    if center_dataset:
                        mean = _centre(sink, mean)
                    ship.flush(0, sink, n_saved_chunks)
                    n_saved_chunks += 1
                    if n_saved_chunks == n_chunks:
                        break
            if n_saved_chunks < n_chunks and sink.fill > 0:  # undersized last chunk (:317-319; not centred there either)
                ship.flush(0, sink, n_saved_chunks)
    finally:
        if handle is not None:
            handle.remove()
        ship.finish()
