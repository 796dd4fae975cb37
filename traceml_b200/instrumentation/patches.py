"""Automatic region openers (mirror of ``src/traceml/instrumentation/patches/*``
and ``hooks/optimizer_hooks.py``).

The *decision* logic is the reference's: thread-local enable flags raised only
inside ``trace_step``, outermost-call-only for forward/backward, forward only
for the traced model (or its DDP / FSDP inner module), CPU->CUDA ``.to()`` only,
one ``dataloader_next`` region per fetched batch, global optimizer pre/post
hooks.  The region *body* is what changed: each opener is a ``timed_region``
whose GPU half is a pair of stamp kernels (``utils/timing.py``).

A single thread-local object carries all flags, so the disabled fast path of
every patched entry point is one attribute read.
"""
from __future__ import annotations

import threading
from typing import Any, Optional

import torch
import torch.nn as nn
from torch.utils.data import DataLoader

from ..utils.timing import timed_region
from .h2d import should_time_h2d


class _Flags(threading.local):
    fwd = False
    fwd_depth = 0
    fwd_targets: frozenset = frozenset()
    bwd = False
    bwd_depth = 0
    h2d = False
    # per-thread reusable region objects: the openers below are depth-guarded (forward,
    # backward) or strictly sequential (h2d), so one object per site is enough and the
    # per-call allocation disappears
    r_fwd = None
    r_bwd = None
    r_h2d = None
    h2d_depth = 0


_TLS = _Flags()

_ORIG_MODULE_CALL = nn.Module.__call__
_ORIG_TENSOR_BACKWARD = torch.Tensor.backward
_ORIG_AUTOGRAD_BACKWARD = torch.autograd.backward
_ORIG_TENSOR_TO = torch.Tensor.to
_ORIG_DATALOADER_ITER = DataLoader.__iter__

FWD, BWD, H2D, DL, OPT = (
    "_traceml_internal:forward_time", "_traceml_internal:backward_time",
    "_traceml_internal:h2d_time", "_traceml_internal:dataloader_next",
    "_traceml_internal:optimizer_step",
)


# ----------------------------------------------------------------- forward
def _module_call(self, *args, **kwargs):
    t = _TLS
    if not t.fwd or t.fwd_depth > 0 or (t.fwd_targets and id(self) not in t.fwd_targets):
        return _ORIG_MODULE_CALL(self, *args, **kwargs)
    t.fwd_depth += 1
    try:
        region = t.r_fwd
        if region is None:
            region = t.r_fwd = timed_region(FWD, "step", True)
        with region:
            return _ORIG_MODULE_CALL(self, *args, **kwargs)
    finally:
        t.fwd_depth -= 1


def patch_forward() -> None:
    if getattr(nn.Module, "_traceml_forward_patched", False):
        return
    nn.Module.__call__ = _module_call  # type: ignore[assignment]
    nn.Module._traceml_forward_patched = True


def forward_targets(model: Optional[nn.Module]) -> frozenset:
    """The model plus its DDP ``.module`` / FSDP ``._fsdp_wrapped_module``
    (forward_auto_timer_patch.py:33-47)."""
    if model is None:
        return frozenset()
    ids = {id(model)}
    for attr in ("module", "_fsdp_wrapped_module"):
        inner = getattr(model, attr, None)
        if isinstance(inner, nn.Module):
            ids.add(id(inner))
    return frozenset(ids)


# ----------------------------------------------------------------- backward
def _tensor_backward(self, *args, **kwargs):
    t = _TLS
    if not t.bwd or t.bwd_depth > 0:
        return _ORIG_TENSOR_BACKWARD(self, *args, **kwargs)
    t.bwd_depth += 1
    try:
        region = t.r_bwd
        if region is None:
            region = t.r_bwd = timed_region(BWD, "step", True)
        with region:
            return _ORIG_TENSOR_BACKWARD(self, *args, **kwargs)
    finally:
        t.bwd_depth -= 1


def _autograd_backward(*args, **kwargs):
    t = _TLS
    if not t.bwd or t.bwd_depth > 0:
        return _ORIG_AUTOGRAD_BACKWARD(*args, **kwargs)
    t.bwd_depth += 1
    try:
        region = t.r_bwd
        if region is None:
            region = t.r_bwd = timed_region(BWD, "step", True)
        with region:
            return _ORIG_AUTOGRAD_BACKWARD(*args, **kwargs)
    finally:
        t.bwd_depth -= 1


def patch_backward() -> None:
    if getattr(torch, "_traceml_backward_patched", False):
        return
    torch.Tensor.backward = _tensor_backward  # type: ignore[assignment]
    torch.autograd.backward = _autograd_backward  # type: ignore[assignment]
    torch._traceml_backward_patched = True  # type: ignore[attr-defined]


# ----------------------------------------------------------------- h2d
def _tensor_to(self, *args, **kwargs):
    t = _TLS
    if not t.h2d or not should_time_h2d(self, args, kwargs):
        return _ORIG_TENSOR_TO(self, *args, **kwargs)
    if t.h2d_depth > 0:
        # a .to() issued inside a timed .to() (tensor subclasses, batch wrappers): the shared
        # region object is open -- time the inner call with its own object
        with timed_region(H2D, "step", True):
            return _ORIG_TENSOR_TO(self, *args, **kwargs)
    region = t.r_h2d
    if region is None:
        region = t.r_h2d = timed_region(H2D, "step", True)
    t.h2d_depth += 1
    try:
        with region:
            return _ORIG_TENSOR_TO(self, *args, **kwargs)
    finally:
        t.h2d_depth -= 1


def patch_h2d() -> None:
    if getattr(torch.Tensor, "_traceml_h2d_patched", False):
        return
    torch.Tensor.to = _tensor_to  # type: ignore[assignment]
    torch.Tensor._traceml_h2d_patched = True  # type: ignore[attr-defined]


# ----------------------------------------------------------------- dataloader
def _dataloader_iter(self):
    it = _ORIG_DATALOADER_ITER(self)
    while True:
        try:
            with timed_region(DL, "step", False):
                batch = next(it)
        except StopIteration:
            return
        yield batch


def patch_dataloader() -> None:
    if getattr(DataLoader, "_traceml_patched", False):
        return
    DataLoader.__iter__ = _dataloader_iter  # type: ignore[assignment]
    DataLoader._traceml_patched = True


# ----------------------------------------------------------------- optimizer
_OPT_OPEN: dict = {}
_OPT_HANDLES = None


def install_optimizer_time_hooks() -> None:
    """Global optimizer pre/post step hooks (hooks/optimizer_hooks.py:17-92)."""
    global _OPT_HANDLES
    if _OPT_HANDLES is not None:
        return
    from torch.optim.optimizer import (register_optimizer_step_post_hook,
                                       register_optimizer_step_pre_hook)

    def pre(optimizer, args, kwargs):
        try:
            region = timed_region(OPT, "step", True)
            region.__enter__()
            _OPT_OPEN[id(optimizer)] = region
        except Exception:
            pass

    def post(optimizer, args, kwargs):
        region = _OPT_OPEN.pop(id(optimizer), None)
        if region is not None:
            try:
                region.__exit__(None, None, None)
            except Exception:
                pass

    _OPT_HANDLES = (register_optimizer_step_pre_hook(pre), register_optimizer_step_post_hook(post))


def ensure_optimizer_timing_installed() -> None:
    if getattr(torch.optim.Optimizer, "_traceml_opt_hooks_installed", False):
        return
    install_optimizer_time_hooks()
    torch.optim.Optimizer._traceml_opt_hooks_installed = True


# ----------------------------------------------------------------- step scope
class step_auto_timers:
    """Raises the forward / backward / h2d enable flags for one ``trace_step``
    (the three nested context managers of sdk/instrumentation.py:179-183 in one)."""

    __slots__ = ("model", "prev")

    def __init__(self, model: Optional[nn.Module]):
        self.model = model
        self.prev = None

    def __enter__(self):
        t = _TLS
        self.prev = (t.fwd, t.fwd_depth, t.fwd_targets)
        t.fwd, t.fwd_depth, t.fwd_targets = True, 0, forward_targets(self.model)
        t.bwd = True
        t.h2d = True
        return self

    def __exit__(self, exc_type, exc, tb):
        t = _TLS
        t.fwd, t.fwd_depth, t.fwd_targets = self.prev
        t.bwd, t.bwd_depth = False, 0
        t.h2d = False
        return False


__all__ = ["patch_forward", "patch_backward", "patch_h2d", "patch_dataloader",
           "install_optimizer_time_hooks", "ensure_optimizer_timing_installed",
           "step_auto_timers", "forward_targets"]
