"""sparse_coding_b200 — B200-native engine for the ensemble sparse-autoencoder sweep of HoagyC/sparse_coding.

Public names mirror the reference's ``autoencoders`` package for the hot path only:
DictSignature / FunctionalEnsemble (ensemble.py), FunctionalSAE / FunctionalTiedSAE / masked variants
(sae_ensemble.py), TopKEncoder / TopKLearnedDict (topk_encoder.py), LearnedDict / TiedSAE / UntiedSAE
(learned_dict.py), plus the driver loop pieces of big_sweep.py (train_loop.py)."""
from .ensemble import CodeProxy, FunctionalEnsemble, optim_str_to_func, stack_dict, unstack_dict
from .learned_dict import LearnedDict, TiedSAE, UntiedSAE
from .optim import AdamConfig, adam
from .sae_ensemble import FunctionalMaskedSAE, FunctionalMaskedTiedSAE, FunctionalSAE, FunctionalTiedSAE
from .signatures import DictSignature
from .topk_encoder import TopKEncoder, TopKLearnedDict

__all__ = [
    "AdamConfig", "CodeProxy", "DictSignature", "FunctionalEnsemble", "FunctionalMaskedSAE", "FunctionalMaskedTiedSAE",
    "FunctionalSAE", "FunctionalTiedSAE", "LearnedDict", "TiedSAE", "TopKEncoder", "TopKLearnedDict", "UntiedSAE",
    "adam", "optim_str_to_func", "stack_dict", "unstack_dict",
]
