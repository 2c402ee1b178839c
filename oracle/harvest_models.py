"""Tiny stand-in language models for the activation-harvesting fixtures.

TEST INFRASTRUCTURE (imported by oracle/make_harvest_golden.py and tests/test_harvest.py only). The reference
harvests from pretrained Pythia / GPT-2 checkpoints through three capture mechanisms, none of which can be
downloaded or installed here:

* HF ``AutoModelForCausalLM`` + forward hooks        (activation_dataset.py:393-497)  -> ``tiny_neox``
* TransformerLens ``HookedTransformer.run_with_cache`` (activation_dataset.py:323-391) -> ``TinyHooked``
* baukit ``Trace`` on a nanoGPT module name            (activation_dataset.py:263-321) -> ``TinyNano``

``TinyHooked`` exposes the two things the reference touches on a HookedTransformer — ``cfg.model_name`` and
``run_with_cache(tokens, stop_at_layer=...) -> (logits, cache)`` with TransformerLens' hook-point names as keys —
on top of a randomly initialised GPT-NeoX (the Pythia architecture). What the harvest code does with the cached
tensors (cast, flatten, chunk, centre, save) is what the fixtures pin; the language model's forward pass is library
code on either side.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Optional

import torch
import torch.nn as nn

TINY_TL_NAME = "tiny-neox"      # the one name the stubbed TransformerLens registry knows
D_MODEL, N_LAYERS, N_HEADS, D_MLP, VOCAB, CTX = 32, 3, 4, 64, 120, 16


def tiny_neox(seed: int = 0):
    import transformers
    cfg = transformers.GPTNeoXConfig(vocab_size=VOCAB, hidden_size=D_MODEL, num_hidden_layers=N_LAYERS,
                                     num_attention_heads=N_HEADS, intermediate_size=D_MLP,
                                     max_position_embeddings=2 * CTX)
    torch.manual_seed(seed)
    return transformers.GPTNeoXForCausalLM(cfg).eval()


class TinyHooked(nn.Module):
    """``run_with_cache`` over a GPT-NeoX with TransformerLens' names for the hook points the reference reads
    (``make_tensor_name``, activation_dataset.py:69-106): ``blocks.{l}.hook_resid_post``, ``blocks.{l}.mlp.hook_post``,
    ``blocks.{l}.hook_mlp_out`` as ``[b, s, n]`` and ``blocks.{l}.attn.hook_z`` as ``[b, s, heads, d_head]``."""

    def __init__(self, lm: Optional[nn.Module] = None):
        super().__init__()
        self.lm = lm if lm is not None else tiny_neox()
        self.cfg = SimpleNamespace(model_name=TINY_TL_NAME, d_model=D_MODEL, d_mlp=D_MLP, n_heads=N_HEADS,
                                   d_head=D_MODEL // N_HEADS, n_layers=N_LAYERS)

    def run_with_cache(self, tokens, stop_at_layer: Optional[int] = None):
        cache: Dict[str, torch.Tensor] = {}
        handles = []
        layers = self.lm.gpt_neox.layers
        upto = len(layers) if stop_at_layer is None else min(stop_at_layer, len(layers))
        for l in range(upto):
            blk = layers[l]
            handles.append(blk.register_forward_hook(
                lambda m, i, o, l=l: cache.__setitem__(f"blocks.{l}.hook_resid_post", o[0] if isinstance(o, tuple) else o)))
            handles.append(blk.mlp.act.register_forward_hook(
                lambda m, i, o, l=l: cache.__setitem__(f"blocks.{l}.mlp.hook_post", o)))
            handles.append(blk.mlp.register_forward_hook(
                lambda m, i, o, l=l: cache.__setitem__(f"blocks.{l}.hook_mlp_out", o)))
            handles.append(blk.attention.dense.register_forward_hook(
                lambda m, i, o, l=l: cache.__setitem__(
                    f"blocks.{l}.attn.hook_z", i[0].reshape(i[0].shape[0], i[0].shape[1], N_HEADS, -1))))
        try:
            with torch.no_grad():
                out = self.lm(tokens)
        finally:
            for h in handles:
                h.remove()
        return out.logits, cache


class _NanoMLP(nn.Module):
    def __init__(self):
        super().__init__()
        self.c_fc = nn.Linear(D_MODEL, D_MLP)
        self.c_proj = nn.Linear(D_MLP, D_MODEL)

    def forward(self, x):
        return self.c_proj(torch.nn.functional.gelu(self.c_fc(x)))


class _NanoBlock(nn.Module):
    def __init__(self):
        super().__init__()
        self.ln = nn.LayerNorm(D_MODEL)
        self.mlp = _NanoMLP()

    def forward(self, x):
        return x + self.mlp(self.ln(x))


class TinyNano(nn.Module):
    """Module tree with nanoGPT's names (``transformer.h.{l}.mlp.c_fc``, the tensor the reference traces with
    baukit, activation_dataset.py:88-89); attention is left out — the harvest only needs a named module output."""

    def __init__(self, seed: int = 0):
        super().__init__()
        torch.manual_seed(seed)
        self.transformer = nn.ModuleDict(dict(wte=nn.Embedding(VOCAB, D_MODEL),
                                              h=nn.ModuleList([_NanoBlock() for _ in range(N_LAYERS)])))
        self.cfg = SimpleNamespace(model_name="nanoGPT")

    def forward(self, tokens):
        x = self.transformer["wte"](tokens)
        for blk in self.transformer["h"]:
            x = blk(x)
        return x
