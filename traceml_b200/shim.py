"""Option B of INTEGRATION.md as code: bind the REFERENCE package's instrumentation seam to
this engine, so that a script that does ``import traceml`` -- the reference's own ``init`` /
``trace_step`` / ``trace_time`` / ``wrap_*`` / patches / Lightning and HF integrations, unmodified
-- records through the C-ABI of ``libtraceml_b200.so`` instead of the CUDA-event pool, the
``TimeEvent`` objects and the Python queues.

    import traceml                      # the reference, unchanged
    import traceml_b200.shim as shim
    shim.install()                      # once, before training starts
    traceml.init(mode="auto")
    with traceml.trace_step(model): ...

What is rebound (SURVEY 8b "instrumentation seam"; reference file:line):
    utils/timing.py:184     timed_region            -> traceml_b200.utils.timing.timed_region
    utils/timing.py:148     record_event            -> ... record_event
    utils/timing.py:163     flush_step_time_buffer  -> no-op (one commit kernel closes the step)
    utils/step_memory.py:30 StepMemoryTracker       -> ... StepMemoryTracker (c10 peaks, no dict)
    utils/step_memory.py:93 flush_step_memory_buffer-> no-op
    utils/flush_buffers.py:24 flush_step_events     -> ... flush_step_events (tml_step_commit)
    hooks/optimizer_hooks.py:17 install_optimizer_time_hooks -> instrumentation.patches.* (stamps)
    samplers/{step_time,step_memory,process}_sampler -> traceml_b200.samplers.* (drain of the ring)
The reference binds these names with ``from ... import name``, so every already-imported
``traceml.*`` module that holds one of them is patched in place; modules imported later pick the
replacements up from the patched home modules.  ``uninstall()`` restores everything.

The reference's step counter (``runtime/state.py``), its patches' decision logic (which call is
timed, under which name) and its public API stay the reference's own code: only the region body,
the step close and the samplers change -- exactly the boundary ``include/traceml_b200.h`` declares.
"""
from __future__ import annotations

import importlib
import sys
from typing import Any, Dict, List, Tuple

_SAVED: List[Tuple[Any, str, Any]] = []
_INSTALLED = False

# seam name -> (home module in the reference, replacement factory)
_SEAM = {
    "timed_region": "traceml.utils.timing",
    "record_event": "traceml.utils.timing",
    "flush_step_time_buffer": "traceml.utils.timing",
    "StepMemoryTracker": "traceml.utils.step_memory",
    "flush_step_memory_buffer": "traceml.utils.step_memory",
    "flush_step_events": "traceml.utils.flush_buffers",
    # the optimizer hooks open their region by hand (two pooled CUDA events + a TimeEvent,
    # hooks/optimizer_hooks.py:17-92): rebind the installer, the hooks become stamp kernels
    "install_optimizer_time_hooks": "traceml.instrumentation.hooks.optimizer_hooks",
    "ensure_optimizer_timing_installed": "traceml.instrumentation.hooks.optimizer_hooks",
}


def _replacements() -> Dict[str, Any]:
    from .instrumentation import patches as pt
    from .utils import flush_buffers as fb
    from .utils import step_memory as sm
    from .utils import timing as tm

    def _noop_flush_memory(model, step):  # utils/step_memory.py:93-110: folded into the commit
        return None

    return {
        "timed_region": tm.timed_region,
        "record_event": tm.record_event,
        "flush_step_time_buffer": tm.flush_step_time_buffer,
        "StepMemoryTracker": sm.StepMemoryTracker,
        "flush_step_memory_buffer": _noop_flush_memory,
        "flush_step_events": fb.flush_step_events,
        "install_optimizer_time_hooks": pt.install_optimizer_time_hooks,
        "ensure_optimizer_timing_installed": pt.ensure_optimizer_timing_installed,
    }


def install(samplers: bool = True) -> List[str]:
    """Bind the reference's seam to this engine.  Returns the patched ``module.name`` list.
    Raises ImportError if the reference package is not importable, RuntimeError if the native
    library is missing (no CPU fallback)."""
    global _INSTALLED
    if _INSTALLED:
        return []
    from . import _abi

    _abi.lib()  # fail loudly before touching anything
    ref = importlib.import_module("traceml")
    if getattr(ref, "__name__", "") != "traceml" or "traceml_b200" in (getattr(ref, "__file__", "") or ""):
        raise ImportError("`traceml` does not resolve to the reference package")
    for mod in set(_SEAM.values()):
        importlib.import_module(mod)
    # the SDK / patch modules bind the names at import: make sure they are loaded before patching
    for mod in ("traceml.sdk.instrumentation", "traceml.sdk.wrappers",
                "traceml.instrumentation.patches.forward_auto_timer_patch",
                "traceml.instrumentation.patches.backward_auto_timer_patch",
                "traceml.instrumentation.patches.h2d_auto_timer_patch",
                "traceml.instrumentation.patches.dataloader_patch",
                "traceml.instrumentation.hooks.optimizer_hooks"):
        try:
            importlib.import_module(mod)
        except Exception:  # optional pieces (an integration's dependency missing) are not the seam
            pass
    new = _replacements()
    originals = {name: getattr(sys.modules[home], name) for name, home in _SEAM.items()}
    patched: List[str] = []
    for modname, mod in list(sys.modules.items()):
        if mod is None or not (modname == "traceml" or modname.startswith("traceml.")):
            continue
        for name, orig in originals.items():
            if getattr(mod, name, None) is orig:
                _SAVED.append((mod, name, orig))
                setattr(mod, name, new[name])
                patched.append(f"{modname}.{name}")
    if samplers:
        patched += _install_samplers()
    _INSTALLED = True
    return sorted(patched)


def _install_samplers() -> List[str]:
    """The ``run``-profile samplers on this path (runtime/sampler_registry.py:78-160) become drains
    of the engine's ring: same class names, ``sampler_name``, table names and ``sample()`` contract."""
    from . import samplers as mine
    from .runtime import get_engine

    out = []
    shared: Dict[str, Any] = {}

    def tap():
        if "tap" not in shared:
            shared["tap"] = mine.RecordTap(get_engine())
        return shared["tap"]

    def make(cls):
        class _Bound(cls):  # zero-argument constructor, like the reference's samplers
            def __init__(self, *a, **k):
                super().__init__(tap())
        _Bound.__name__ = cls.__name__
        _Bound.__qualname__ = cls.__qualname__
        return _Bound

    for modname, name, cls in (("traceml.samplers.step_time_sampler", "StepTimeSampler", mine.StepTimeSampler),
                               ("traceml.samplers.step_memory_sampler", "StepMemorySampler", mine.StepMemorySampler),
                               ("traceml.samplers.process_sampler", "ProcessSampler", mine.ProcessSampler)):
        try:
            mod = importlib.import_module(modname)
        except Exception:
            continue
        orig = getattr(mod, name)
        bound = make(cls)
        for mname, m in list(sys.modules.items()):
            if m is not None and mname.startswith("traceml.") and getattr(m, name, None) is orig:
                _SAVED.append((m, name, orig))
                setattr(m, name, bound)
                out.append(f"{mname}.{name}")
    return out


def uninstall() -> None:
    global _INSTALLED
    while _SAVED:
        mod, name, orig = _SAVED.pop()
        setattr(mod, name, orig)
    _INSTALLED = False


def installed() -> bool:
    return _INSTALLED


__all__ = ["install", "uninstall", "installed"]
