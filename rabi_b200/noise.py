"""Base-noise source for the MC estimators.

By default standard-normal noise is drawn on the device with torch's CUDA generator.  The reference draws it on
whatever device the model lives on through ``Normal.rsample`` (one ``normal_`` of shape [S, *batch, D] per
``rsample``); CPU and CUDA generators give different streams, so result parity is defined on
*injected* noise: inside ``with injected([...])`` the tensors are served in call order instead of fresh draws.
"""
from __future__ import annotations

import contextlib
from typing import List, Sequence

import torch

_queue: List[torch.Tensor] | None = None


def standard_normal(shape, device) -> torch.Tensor:
    """fp32 N(0,1) tensor of ``shape`` on ``device`` (or the next injected tensor, which must match in numel)."""
    global _queue
    if _queue is not None:
        if not _queue:
            raise RuntimeError("injected noise exhausted")
        z = _queue.pop(0)
        shape = tuple(shape)
        if z.numel() != int(torch.Size(shape).numel()):
            raise RuntimeError(f"injected noise has shape {tuple(z.shape)}, the estimator needs {shape}")
        return z.to(device=device, dtype=torch.float32).reshape(shape).contiguous()
    return torch.randn(tuple(shape), device=device, dtype=torch.float32)


@contextlib.contextmanager
def injected(noise: Sequence[torch.Tensor]):
    global _queue
    prev = _queue
    _queue = list(noise)
    try:
        yield _queue
    finally:
        _queue = prev
