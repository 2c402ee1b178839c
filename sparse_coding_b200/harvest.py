"""Activation harvesting — the producer of the `{folder}/{tensor_name}/{i}.pt` fp16 chunks the sweep trains on
(reference: ``activation_dataset.py:393-503`` ``make_activation_dataset_hf`` / ``save_activation_chunk``).

The reference's hook casts every layer output to fp16 and moves it to the host *inside the forward pass*
(``.to(dtype).cpu()``: one synchronous D2H copy per hooked module per model batch), keeps a Python list of those
pieces and concatenates them on the CPU at every chunk boundary. Here the hook only copies (with the cast fused
into the copy) into a preallocated device buffer of one chunk; a full chunk leaves the device in a single
asynchronous copy to pinned memory on a side stream while the language model keeps running into the second buffer,
and a background thread writes the file. The on-disk format is unchanged: ``torch.save`` of one ``[rows, d]``
fp16 (or fp32) tensor per chunk. The language model itself is whatever ``torch.nn.Module`` the caller passes
(the reference uses HF ``AutoModelForCausalLM``); its forward pass is library code, not part of this engine.

Deliberate difference: the reference's chunk-boundary test ``batch_idx+1 % chunk_batches == 0`` parses as
``batch_idx + (1 % chunk_batches) == 0`` and therefore never fires (everything lands in one "undersized final
chunk"); chunks here are cut every ``chunk_size // (model_batch_size * max_length)`` batches as intended.
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
    """Two device buffers of one chunk each for one hooked tensor; rows are appended by the forward hook."""

    def __init__(self, rows: int, dtype: torch.dtype, device: torch.device, folder: str):
        self.rows, self.dtype, self.device, self.folder = rows, dtype, device, folder
        self.bufs: List[Optional[torch.Tensor]] = [None, None]
        self.pinned: List[Optional[torch.Tensor]] = [None, None]
        self.copied = [None, None]          # event: D2H out of buffer `slot` finished
        self.writer = [None, None]          # future: the file write that reads pinned buffer `slot`
        self.slot, self.fill, self.width = 0, 0, None

    def append(self, out: torch.Tensor, copy_stream) -> None:
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
        self.bufs[self.slot][self.fill:self.fill + take].copy_(flat[:take])          # cast fused into the copy
        self.fill += take
        # rows beyond the chunk boundary are dropped, as a chunk is a whole number of model batches

    def full(self) -> bool:
        return self.fill >= self.rows


def make_activation_dataset_hf(sentence_dataset, model: torch.nn.Module, tensor_names: List[str], chunk_size: int,
                               n_chunks: int, output_folder: str = "activation_data", skip_chunks: int = 0,
                               device: Optional[torch.device] = torch.device("cuda:0"), max_length: int = 2048,
                               model_batch_size: int = 4, precision: str = "float16",
                               shuffle_seed: Optional[int] = None) -> Dict[str, List[str]]:
    """Run ``model`` over ``sentence_dataset`` (items with an ``"input_ids"`` tensor of ``max_length`` tokens) and
    write, for every module name in ``tensor_names``, chunks of ``chunk_size`` activation rows to
    ``{output_folder}/{tensor_name}/{i}.pt``. Same signature and file layout as the reference
    (activation_dataset.py:393-405). Returns the written paths per tensor name."""
    if precision == "float16":
        dtype = torch.float16
    elif precision == "float32":
        dtype = torch.float32
    else:
        raise ValueError(f"Invalid precision '{precision}'")
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

    on_gpu = device.type == "cuda"
    copy_stream = torch.cuda.Stream(device) if on_gpu else None
    sinks = {name: _ChunkSink(rows_per_chunk, dtype, device, os.path.join(output_folder, name)) for name in tensor_names}
    written: Dict[str, List[str]] = {name: [] for name in tensor_names}
    pool = ThreadPoolExecutor(max_workers=1)
    pending = []
    handles = []
    modules = dict(model.named_modules())
    for name in tensor_names:
        if name not in modules:
            raise KeyError(f"module '{name}' not found in the model")

        def hook(module, inputs, output, name=name):
            out = output[0] if isinstance(output, tuple) else output
            sinks[name].append(out.detach(), copy_stream)
            return output

        handles.append(modules[name].register_forward_hook(hook))

    def flush(name: str, chunk_idx: int, rows: int):
        """Ship buffer `slot` of this sink: async D2H into pinned memory, file write on the worker thread."""
        sink = sinks[name]
        slot = sink.slot
        src = sink.bufs[slot][:rows]
        if on_gpu:
            if sink.pinned[slot] is None:
                sink.pinned[slot] = torch.empty(sink.rows, sink.width, dtype=dtype).pin_memory()
            if sink.writer[slot] is not None:
                sink.writer[slot].result()                                   # the previous file out of this pinned buffer is written
            copy_stream.wait_stream(torch.cuda.current_stream(device))       # the hooks' copies are done
            with torch.cuda.stream(copy_stream):
                sink.pinned[slot][:rows].copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            sink.copied[slot] = ev
            host = sink.pinned[slot][:rows]

            def job(ev=ev, host=host, full=rows == sink.rows):
                ev.synchronize()
                # torch.save serialises the whole underlying storage of a view: an undersized final chunk is written
                # from a compact copy (the reference writes a compact torch.cat result, activation_dataset.py:499-503)
                return save_activation_chunk(host if full else host.clone(), chunk_idx, sink.folder)
        else:
            host = src.clone()

            def job(host=host):
                return save_activation_chunk(host, chunk_idx, sink.folder)
        fut = pool.submit(job)
        sink.writer[slot] = fut
        pending.append((name, fut))
        sink.slot ^= 1
        sink.fill = 0

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
                        flush(name, chunk_idx, sinks[name].fill)
                    chunk_idx += 1
                    batches_in_chunk = 0
                    if chunk_idx >= n_chunks:
                        break
            if chunk_idx < n_chunks and batches_in_chunk > 0:           # undersized final chunk
                for name in tensor_names:
                    flush(name, chunk_idx, sinks[name].fill)
    finally:
        for h in handles:
            h.remove()
        for name, fut in pending:
            written[name].append(fut.result())
        pool.shutdown()
    return written
