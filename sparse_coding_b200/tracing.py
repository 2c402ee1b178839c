"""NVTX ranges around the host-visible phases of a run (SURVEY §5 "tracing"): ``SCE_NVTX=1`` names the step calls, the
chunk staging and the checkpoint writes in an Nsight Systems / ncu ``--nvtx`` timeline. Off by default: a range is two
driver calls per step, about 2 % of a config-1 step (80 us)."""
from __future__ import annotations

import contextlib
import os

import torch

ENABLED = os.environ.get("SCE_NVTX", "0") == "1"


@contextlib.contextmanager
def _on(name: str):
    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()


_OFF = contextlib.nullcontext()


def nvtx_range(name: str):
    """Context manager: an NVTX range called ``name`` when tracing is enabled, else nothing."""
    return _on(name) if ENABLED else _OFF
