"""Optimiser description for the engine.

The reference passes ``torchopt.adam`` and ``{"lr": …}`` into FunctionalEnsemble (basic_l1_sweep.py:69-73,
big_sweep_experiments.py:69-79) and lets ``vmap(optimizer.update)`` run ~12 elementwise kernels per leaf. In the
engine the whole update of a dictionary row — row-norm Jacobian of the weight gradient, Adam moments and step,
re-normalisation and the re-split into operand planes for the next step's GEMMs — is ONE streaming kernel that runs
after the weight-gradient GEMM (csrc/sce_kernels.cuh: dict_rows_kernel<MODE_ADAM>; DESIGN.md §4 explains why it is
not that GEMM's epilogue), so on the Python side an optimiser is just its hyper-parameters. ``adam`` mirrors torchopt.adam's keyword names
(``lr, betas, eps, eps_root``; weight decay and the other torchopt options are not supported and raise)."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class AdamConfig:
    lr: float = 1e-3
    b1: float = 0.9
    b2: float = 0.999
    eps: float = 1e-8
    eps_root: float = 0.0


def adam(lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, *, eps_root: float = 0.0,
         **unsupported) -> AdamConfig:
    if weight_decay:
        raise NotImplementedError("weight_decay is not on the reference hot path (callers pass only lr)")
    extra = {k: v for k, v in unsupported.items() if v not in (None, False)}
    if extra:
        raise NotImplementedError(f"unsupported torchopt.adam options for the fused engine: {sorted(extra)}")
    return AdamConfig(float(lr), float(betas[0]), float(betas[1]), float(eps), float(eps_root))


def resolve_optimizer(optimizer_func, optimizer_kwargs) -> AdamConfig:
    """Accepts ``adam`` (this module), the string "adam", or ``torchopt.adam`` when torchopt is installed."""
    kwargs = dict(optimizer_kwargs or {})
    if isinstance(optimizer_func, str):
        if optimizer_func != "adam":
            raise ValueError("Unknown optimizer string: {}".format(optimizer_func))
        return adam(**kwargs)
    if optimizer_func is adam:
        return adam(**kwargs)
    name = getattr(optimizer_func, "__name__", "")
    module = getattr(optimizer_func, "__module__", "") or ""
    if name.lower() == "adam" and module.split(".")[0] == "torchopt":
        return adam(**kwargs)
    raise ValueError(f"the fused engine implements Adam only; got optimizer_func={optimizer_func!r}")
