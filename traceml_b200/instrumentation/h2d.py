"""Which ``.to(...)`` calls count as host-to-device input movement
(mirror of ``src/traceml/instrumentation/h2d.py:46-67``): CUDA target, source
not already on CUDA, not an ``nn.Parameter`` (so ``model.to()`` is ignored)."""
from __future__ import annotations

from typing import Any, Optional

import torch


def _target_type(value: Any) -> Optional[str]:
    if isinstance(value, torch.device):
        return value.type
    if isinstance(value, torch.Tensor):
        return value.device.type
    if isinstance(value, str):
        try:
            return torch.device(value).type
        except (RuntimeError, TypeError):
            return None
    return None


def is_cuda_target(args, kwargs) -> bool:
    if args and _target_type(args[0]) == "cuda":
        return True
    return _target_type(kwargs.get("device")) == "cuda"


def should_time_h2d(obj: Any, args, kwargs) -> bool:
    if not is_cuda_target(args, kwargs):
        return False
    if isinstance(obj, torch.nn.Parameter):
        return False
    if isinstance(obj, torch.Tensor):
        return not obj.is_cuda
    return True


__all__ = ["should_time_h2d", "is_cuda_target"]
