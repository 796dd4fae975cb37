"""Step flush seam (mirror of ``src/traceml/utils/flush_buffers.py:24-33``).

One call closes the step: the commit kernel merges the device-side phase
accumulators with the host-clock phases and the allocator peaks into one 128-B
StepRecord in the HBM ring.  No host synchronisation.
"""
from __future__ import annotations

import sys
import time

from ..records import FLAG_HAS_MEM
from ..runtime import disabled, get_engine
from .step_memory import take_pending


def flush_step_events(model, step: int) -> None:
    if disabled():
        return
    from . import timing

    try:
        eng = timing._ENG or timing._resolve()
        pend = take_pending(model)
        alloc = resv = 0
        flags = 0
        if pend is not None and pend[0] is not None:
            alloc, resv, flags = pend[0], pend[1], FLAG_HAS_MEM
        rc = eng._commit(eng._h, int(step), alloc, resv, flags, time.time(),
                         timing._raw_stream(timing._cur_dev()))
        if rc < 0:
            eng.step_discard()
            print(f"[TraceML] step {step} not committed (status {rc})", file=sys.stderr)
        _layers = sys.modules.get("traceml_b200.instrumentation.layers")
        if _layers is not None and _layers._PROFILES:  # deep profile: close the step for every layer too
            _layers.commit_all(int(step))
    except timing._Quiet:
        pass  # engine resolution failed and was reported once
    except Exception as exc:
        print(f"[TraceML] flush failed: {exc}", file=sys.stderr)


__all__ = ["flush_step_events"]
