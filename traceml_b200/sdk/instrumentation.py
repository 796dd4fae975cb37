"""``trace_step`` (mirror of ``src/traceml/sdk/instrumentation.py:160-200``).

Step boundary semantics kept exactly: reset allocator peaks; open the
host-clock ``step_time`` region; raise the auto-timer flags; install optimizer
hooks on first use in auto mode; on exit advance the step counter only if the
body completed, record peaks, and flush under the (possibly advanced) step id
-- a failed step therefore flushes under the old id.  Telemetry never raises;
user exceptions propagate.
"""
from __future__ import annotations

import functools
import os
import sys
from typing import Callable

from ..runtime import disabled, get_trace_session_state
from ..utils.flush_buffers import flush_step_events
from ..utils.step_memory import StepMemoryTracker
from ..utils.timing import timed_region

STEP = "_traceml_internal:step_time"


def _auto_optimizer() -> bool:
    from .initial import get_init_config

    cfg = get_init_config()
    return cfg is None or cfg.mode == "auto"


def _log(message: str, exc: Exception) -> None:
    print(f"[TraceML] {message}: {exc}", file=sys.stderr)


class trace_step:
    """``with trace_step(model): ...`` -- one training step."""

    __slots__ = ("model", "tracker", "region", "flags", "active")

    def __init__(self, model):
        self.model = model
        self.active = not disabled()
        self.tracker = None
        self.region = None
        self.flags = None

    def __enter__(self):
        if not self.active:
            return self
        from ..instrumentation import patches

        try:
            self.tracker = StepMemoryTracker(self.model)
            self.tracker.reset()
        except Exception as exc:
            _log("reset failed", exc)
        self.region = timed_region(STEP, "step", False)
        self.region.__enter__()
        self.flags = patches.step_auto_timers(self.model)
        self.flags.__enter__()
        try:
            if _auto_optimizer():
                patches.ensure_optimizer_timing_installed()
        except Exception as exc:
            _log("optimizer hook install failed", exc)
        return self

    def __exit__(self, exc_type, exc, tb):
        if not self.active:
            return False
        self.flags.__exit__(exc_type, exc, tb)
        self.region.__exit__(exc_type, exc, tb)
        state = get_trace_session_state()
        if exc_type is None:
            state.advance_step()
        try:
            if self.tracker is not None:
                self.tracker.record()
        except Exception as e:
            _log("record failed", e)
        try:
            flush_step_events(self.model, state.step)
        except Exception as e:
            _log("flush failed", e)
        return False


def trace_model_instance(model, **kwargs) -> None:
    """Deep (per-layer) profile hooks: out of the hot-path scope;
    accepted and ignored unless TRACEML_PROFILE=deep, where it reports that."""
    if disabled() or (os.environ.get("TRACEML_PROFILE", "run") or "run").strip().lower() != "deep":
        return
    print("[TraceML] deep (per-layer) profile is not part of the B200 engine", file=sys.stderr)


def trace_time(name: str, scope: str = "global", use_gpu: bool = True) -> Callable:
    if disabled():
        return lambda func: func
    if scope not in ("step", "global"):
        raise ValueError(f"Invalid scope {scope!r}. Expected 'step' or 'global'.")

    def decorator(func: Callable):
        @functools.wraps(func)
        def wrapper(*args, **kwargs):
            with timed_region(name, scope=scope, use_gpu=use_gpu):
                return func(*args, **kwargs)

        return wrapper

    return decorator


class _TraceStateMeta(type):
    @property
    def step(cls) -> int:
        return get_trace_session_state().step

    @step.setter
    def step(cls, value: int) -> None:
        get_trace_session_state().set_step(value)


class TraceState(metaclass=_TraceStateMeta):
    """Compatibility facade (``TraceState.step += 1`` keeps working)."""

    @classmethod
    def reset(cls, step: int = 0) -> int:
        return get_trace_session_state().reset(step)

    @classmethod
    def advance(cls, delta: int = 1) -> int:
        return get_trace_session_state().advance_step(delta)


__all__ = ["trace_step", "trace_model_instance", "trace_time", "TraceState"]
