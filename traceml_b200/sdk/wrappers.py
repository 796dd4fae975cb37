"""Manual instrumentation wrappers (mirror of ``src/traceml/sdk/wrappers.py:142-357``).

Each wrapper opens the same canonical region as its automatic twin and refuses
to stack on top of it (duplicate instrumentation raises ``RuntimeError``).
"""
from __future__ import annotations

import functools
from typing import Any

import torch
import torch.nn as nn
from torch.utils.data import DataLoader

from ..instrumentation.h2d import should_time_h2d
from ..instrumentation.patches import BWD, DL, FWD, H2D, OPT
from ..utils.timing import timed_region


def _dup(feature: str, reason: str) -> None:
    raise RuntimeError(
        f"TraceML cannot apply manual wrapper instrumentation for {feature} because automatic "
        f"instrumentation is already active. {reason} Disable the automatic path for this "
        "feature before using the wrapper.")


class _FetchIter:
    def __init__(self, it):
        self._it = it

    def __iter__(self):
        return self

    def __next__(self):
        with timed_region(DL, "step", False):
            return next(self._it)


class _FetchLoader:
    def __init__(self, loader):
        self._loader = loader

    def __iter__(self):
        return _FetchIter(iter(self._loader))

    def __len__(self):
        return len(self._loader)

    def __getattr__(self, name):
        return getattr(self._loader, name)


def wrap_dataloader_fetch(obj: Any) -> Any:
    if getattr(DataLoader, "_traceml_patched", False):
        typ = type(obj)
        is_torch_iter = (str(getattr(typ, "__module__", "")).startswith("torch.utils.data")
                         and str(getattr(typ, "__name__", "")).endswith("DataLoaderIter"))
        if isinstance(obj, DataLoader) or is_torch_iter:
            _dup("dataloader fetch", "torch DataLoader fetch timing is already patched.")
    if hasattr(obj, "__next__"):
        return _FetchIter(obj)
    if hasattr(obj, "__iter__"):
        return _FetchLoader(obj)
    raise TypeError("wrap_dataloader_fetch() expects a loader or iterator object.")


def wrap_forward(model: nn.Module) -> nn.Module:
    if getattr(nn.Module, "_traceml_forward_patched", False):
        _dup("forward", "nn.Module.__call__ has already been patched.")
    if not isinstance(model, nn.Module):
        raise TypeError("wrap_forward() expects an nn.Module instance.")
    if getattr(model, "_traceml_forward_instance_wrapped", False):
        return model
    original = getattr(model, "forward", None)
    if original is None or not callable(original):
        raise TypeError("wrap_forward() requires a callable model.forward.")

    @functools.wraps(original)
    def forward(*args, **kwargs):
        with timed_region(FWD, "step", True):
            return original(*args, **kwargs)

    model.forward = forward  # type: ignore[method-assign]
    model._traceml_forward_instance_wrapped = True
    model._traceml_original_forward = original
    return model


class _BackwardHandle:
    def __init__(self, loss):
        self._loss = loss

    def backward(self, *args, **kwargs):
        with timed_region(BWD, "step", True):
            return self._loss.backward(*args, **kwargs)

    def __getattr__(self, name):
        return getattr(self._loss, name)


def wrap_backward(loss: Any) -> Any:
    if getattr(torch, "_traceml_backward_patched", False):
        _dup("backward", "torch backward entry points have already been patched.")
    if not callable(getattr(loss, "backward", None)):
        raise TypeError("wrap_backward() expects an object with a callable backward() method.")
    return _BackwardHandle(loss)


def wrap_optimizer(optimizer: Any) -> Any:
    if getattr(torch.optim.Optimizer, "_traceml_opt_hooks_installed", False):
        _dup("optimizer step", "global optimizer step hooks are already installed.")
    step_fn = getattr(optimizer, "step", None)
    if step_fn is None or not callable(step_fn):
        raise TypeError("wrap_optimizer() expects an object with a callable step() method.")
    if getattr(optimizer, "_traceml_step_instance_wrapped", False):
        return optimizer

    @functools.wraps(step_fn)
    def step(*args, **kwargs):
        with timed_region(OPT, "step", True):
            return step_fn(*args, **kwargs)

    optimizer.step = step  # type: ignore[method-assign]
    optimizer._traceml_step_instance_wrapped = True
    optimizer._traceml_original_step = step_fn
    return optimizer


class _H2D:
    def __init__(self, obj):
        self._obj = obj

    def to(self, *args, **kwargs):
        if getattr(torch.Tensor, "_traceml_h2d_patched", False) or \
                not should_time_h2d(self._obj, args, kwargs):
            return self._obj.to(*args, **kwargs)
        with timed_region(H2D, "step", True):
            return self._obj.to(*args, **kwargs)

    def __getattr__(self, name):
        return getattr(self._obj, name)

    def __len__(self):
        return len(self._obj)

    def __iter__(self):
        return iter(self._obj)

    def __getitem__(self, key):
        return self._obj[key]

    def __contains__(self, item):
        return item in self._obj

    def __repr__(self):
        return f"_WrappedH2D({self._obj!r})"


def wrap_h2d(obj: Any) -> Any:
    if not callable(getattr(obj, "to", None)):
        raise TypeError("wrap_h2d() expects an object with a callable .to() method.")
    return _H2D(obj)


__all__ = ["wrap_dataloader_fetch", "wrap_forward", "wrap_backward", "wrap_optimizer", "wrap_h2d"]
