"""Policy: which ``.to(...)`` calls are host-to-device *input* movement.

Same decisions as the reference's ``instrumentation/h2d.py:46-67`` (they are part of the
drop-in contract: the h2d phase must count the same calls):

  counted      CPU tensor / user batch wrapper -> a CUDA destination
  not counted  CPU-only moves, dtype-only casts, D2H, D2D, and ``nn.Parameter`` movement
               (so ``model.to("cuda")`` never shows up as input traffic)

This sits on the ``Tensor.to`` patch, i.e. on every ``.to`` of a training step, so the
destination lookup for string arguments is memoised.
"""
from __future__ import annotations

from functools import lru_cache
from typing import Any, Mapping, Sequence

import torch

_CUDA = "cuda"


@lru_cache(maxsize=64)
def _kind_of_name(name: str) -> str:
    try:
        return torch.device(name).type
    except (RuntimeError, TypeError):
        return ""


def _destination_kind(spec: Any) -> str:
    """Device type a ``.to`` argument names ("" when it names none: dtypes, None, ...)."""
    if type(spec) is str:
        return _kind_of_name(spec)
    if isinstance(spec, torch.device):
        return spec.type
    return spec.device.type if isinstance(spec, torch.Tensor) else ""


def is_cuda_target(args: Sequence[Any], kwargs: Mapping[str, Any]) -> bool:
    """``x.to("cuda")``, ``x.to(torch.device("cuda"))``, ``x.to(device=...)``, ``x.to(cuda_tensor)``."""
    positional = _destination_kind(args[0]) if args else ""
    return positional == _CUDA or _destination_kind(kwargs.get("device")) == _CUDA


def should_time_h2d(obj: Any, args: Sequence[Any], kwargs: Mapping[str, Any]) -> bool:
    if isinstance(obj, torch.nn.Parameter) or not is_cuda_target(args, kwargs):
        return False
    # plain tensors count only when they start on the host; an opaque batch wrapper the user
    # chose to wrap counts as input movement (its members cannot be inspected safely)
    return not obj.is_cuda if isinstance(obj, torch.Tensor) else True


__all__ = ["should_time_h2d", "is_cuda_target"]
