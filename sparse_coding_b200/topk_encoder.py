"""TopK dictionary: signature (training side) and learned-dict (inference side).

Mirrors ``autoencoders/topk_encoder.py``: ``TopKEncoder.init`` (:10-17, **randn** dictionary on the CPU, long
``sparsity`` buffer), ``encode`` (:19-27, top-k by SIGNED score, then ReLU), ``to_learned_dict`` (:43-46) and
``TopKLearnedDict`` (:49-62). ``loss`` is executed by the CUDA engine (sparse_coding_b200.ensemble)."""
from __future__ import annotations

import torch

from .learned_dict import LearnedDict
from .signatures import DictSignature, engine_loss

_REF_MODULE = "autoencoders.topk_encoder"


def topk_relu_code(scores: torch.Tensor, k: int) -> torch.Tensor:
    idx = torch.topk(scores, k, dim=-1).indices
    code = torch.zeros_like(scores)
    code.scatter_(-1, idx, scores.gather(-1, idx))
    return code.clamp(min=0.0)


class TopKEncoder(DictSignature):
    variant = "topk"

    @staticmethod
    def init(d_activation, n_features, sparsity, dtype=torch.float32):
        params = {"dict": torch.randn(n_features, d_activation, dtype=dtype)}
        buffers = {"sparsity": torch.tensor(sparsity, dtype=torch.long)}
        return params, buffers

    @staticmethod
    def encode(b, sparsity, normed_dict):
        return topk_relu_code(b @ normed_dict.T, int(sparsity))

    @staticmethod
    def loss(params, buffers, batch):
        return engine_loss(TopKEncoder, params, buffers, batch)

    @staticmethod
    def to_learned_dict(params, buffers):
        d = params["dict"]
        return TopKLearnedDict(d / d.norm(dim=-1)[:, None], buffers["sparsity"].item())


class TopKLearnedDict(LearnedDict):
    def __init__(self, dict, sparsity):
        self.dict = dict
        self.sparsity = sparsity
        self.n_feats, self.activation_size = self.dict.shape

    def to_device(self, device):
        self.dict = self.dict.to(device)

    def encode(self, x):
        return TopKEncoder.encode(x, self.sparsity, self.dict)

    def get_learned_dict(self):
        return self.dict


for _cls in (TopKEncoder, TopKLearnedDict):
    _cls.__module__ = _REF_MODULE
