"""SAE signatures on the accelerated hot path (reference: autoencoders/sae_ensemble.py).

    FunctionalSAE            untied encoder/decoder        sae_ensemble.py:13-78
    FunctionalTiedSAE        tied, optional centring       sae_ensemble.py:81-162
    FunctionalMaskedTiedSAE  tied, per-model dict size     sae_ensemble.py:309-373
    FunctionalMaskedSAE      untied, per-model dict size   sae_ensemble.py:377-444

``init`` keeps the reference's positional orders and initialisers (xavier-uniform matrices, zero bias, 0-dim
hyper-parameter buffers) so that seeded initialisation is bit-identical; ``to_learned_dict`` builds the same export
objects. ``loss`` is evaluated by the CUDA engine. Differences from the reference, both deliberate:
  * FunctionalTiedSAE.init stores ``bias_decay`` in the buffers. The reference accepts the argument, drops it, and
    then reads ``buffers["bias_decay"]`` in ``loss`` (:150) — a KeyError at HEAD.
  * the centring of FunctionalTiedSAE is applied once to the batch, the identity case is skipped, and the unused
    un-centred reconstruction (:146) is not computed.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .learned_dict import TiedSAE, UntiedSAE
from .signatures import DictSignature, engine_loss

_REF_MODULE = "autoencoders.sae_ensemble"


def _xavier(n, d, device, dtype):
    w = torch.empty((n, d), device=device, dtype=dtype)
    nn.init.xavier_uniform_(w)
    return w


class FunctionalSAE(DictSignature):
    variant = "untied"

    @staticmethod
    def init(activation_size, n_dict_components, l1_alpha, bias_decay=0.0, device=None, dtype=None):
        params = {}
        params["encoder"] = _xavier(n_dict_components, activation_size, device, dtype)
        params["encoder_bias"] = torch.zeros((n_dict_components,), device=device, dtype=dtype)
        params["decoder"] = _xavier(n_dict_components, activation_size, device, dtype)
        buffers = {
            "l1_alpha": torch.tensor(l1_alpha, device=device, dtype=dtype),
            "bias_decay": torch.tensor(bias_decay, device=device, dtype=dtype),
        }
        return params, buffers

    @staticmethod
    def to_learned_dict(params, buffers):
        return UntiedSAE(params["encoder"], params["decoder"], params["encoder_bias"])

    @staticmethod
    def encode(params, buffers, batch):
        return (batch @ params["encoder"].T + params["encoder_bias"]).clamp(min=0.0)

    @staticmethod
    def loss(params, buffers, batch):
        return engine_loss(FunctionalSAE, params, buffers, batch)


class FunctionalTiedSAE(DictSignature):
    variant = "tied"

    @staticmethod
    def init(activation_size, n_dict_components, l1_alpha, device=None, dtype=None, bias_decay=0.0,
             translation=None, rotation=None, scaling=None):
        buffers = {}
        buffers["center_rot"] = (torch.eye(activation_size, device=device, dtype=dtype)
                                 if rotation is None else rotation)
        buffers["center_trans"] = (torch.zeros(activation_size, device=device, dtype=dtype)
                                   if translation is None else translation)
        buffers["center_scale"] = (torch.ones(activation_size, device=device, dtype=dtype)
                                   if scaling is None else scaling)
        params = {}
        params["encoder"] = _xavier(n_dict_components, activation_size, device, dtype)
        params["encoder_bias"] = torch.zeros((n_dict_components,), device=device, dtype=dtype)
        buffers["l1_alpha"] = torch.tensor(l1_alpha, device=device, dtype=dtype)
        buffers["bias_decay"] = torch.tensor(bias_decay, device=device, dtype=dtype)  # Q1, see module docstring
        return params, buffers

    @staticmethod
    def to_learned_dict(params, buffers):
        return TiedSAE(params["encoder"], params["encoder_bias"],
                       centering=(buffers["center_trans"], buffers["center_rot"], buffers["center_scale"]),
                       norm_encoder=True)

    @staticmethod
    def center(buffers, batch):
        return ((batch - buffers["center_trans"][None, :]) @ buffers["center_rot"].T) * buffers["center_scale"][None, :]

    @staticmethod
    def uncenter(buffers, batch):
        return (batch / buffers["center_scale"][None, :]) @ buffers["center_rot"] + buffers["center_trans"][None, :]

    @staticmethod
    def loss(params, buffers, batch):
        return engine_loss(FunctionalTiedSAE, params, buffers, batch)


def _mask_buffers(n_dict_components, n_components_stack, l1_alpha, bias_decay, device, dtype):
    mask = torch.ones(n_components_stack, device=device, dtype=torch.bool)
    mask[:n_dict_components] = False
    return {
        "l1_alpha": torch.tensor(l1_alpha, device=device, dtype=dtype),
        "bias_decay": torch.tensor(bias_decay, device=device, dtype=dtype),
        "dict_size": torch.tensor(n_dict_components, device=device, dtype=torch.long),
        "coef_mask": mask,
    }


class FunctionalMaskedTiedSAE(DictSignature):
    """Models with different dictionary sizes share one [M, n_stack, d] stack; coefficients beyond a model's
    ``dict_size`` are forced to zero (sae_ensemble.py:331-333, :356). The bias-decay term is not part of this
    signature's loss (:347-373)."""
    variant = "masked_tied"

    @staticmethod
    def init(activation_size, n_dict_components, n_components_stack, l1_alpha, bias_decay=0.0, device=None,
             dtype=None):
        params = {
            "encoder": _xavier(n_components_stack, activation_size, device, dtype),
            "encoder_bias": torch.zeros((n_components_stack,), device=device, dtype=dtype),
        }
        return params, _mask_buffers(n_dict_components, n_components_stack, l1_alpha, bias_decay, device, dtype)

    @staticmethod
    def to_learned_dict(params, buffers):
        k = buffers["dict_size"].item()
        return TiedSAE(params["encoder"][:k], params["encoder_bias"][:k], norm_encoder=True)

    @staticmethod
    def loss(params, buffers, batch):
        return engine_loss(FunctionalMaskedTiedSAE, params, buffers, batch)


class FunctionalMaskedSAE(DictSignature):
    """Untied counterpart (sae_ensemble.py:377-444)."""
    variant = "masked_untied"

    @staticmethod
    def init(activation_size, n_dict_components, n_components_stack, l1_alpha, bias_decay=0.0, device=None,
             dtype=None):
        params = {
            "encoder": _xavier(n_components_stack, activation_size, device, dtype),
            "encoder_bias": torch.zeros((n_components_stack,), device=device, dtype=dtype),
            "decoder": _xavier(n_components_stack, activation_size, device, dtype),
        }
        return params, _mask_buffers(n_dict_components, n_components_stack, l1_alpha, bias_decay, device, dtype)

    @staticmethod
    def to_learned_dict(params, buffers):
        k = buffers["dict_size"].item()
        return UntiedSAE(params["encoder"][:k], params["decoder"][:k], params["encoder_bias"][:k])

    @staticmethod
    def loss(params, buffers, batch):
        return engine_loss(FunctionalMaskedSAE, params, buffers, batch)


for _cls in (FunctionalSAE, FunctionalTiedSAE, FunctionalMaskedTiedSAE, FunctionalMaskedSAE):
    _cls.__module__ = _REF_MODULE
