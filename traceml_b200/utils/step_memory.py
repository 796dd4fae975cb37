"""Per-step peak-memory seam (mirror of ``src/traceml/utils/step_memory.py:30-110``).

``reset()`` at step start and ``record()`` at step end read c10's caching
allocator counters through the in-tree extension (two integer reads, no Python
dict); ``flush_step_events`` hands the pair to the commit kernel via the
host-mapped counter page.  On a non-CUDA model the peaks are ``None`` and the
step record carries no memory flag (the reference stores NULLs).
"""
from __future__ import annotations

import sys
from typing import Dict, Optional, Tuple

from ..runtime import disabled

_alloc = None
_alloc_tried = False
_pending: Dict[int, Tuple[Optional[int], Optional[int], str]] = {}


def _alloc_ext():
    global _alloc, _alloc_tried
    if not _alloc_tried:
        _alloc_tried = True
        try:
            from .. import _tml_step as _tml_alloc  # built by __graft_entry__.build()

            _alloc = _tml_alloc
        except Exception as exc:
            print(f"[TraceML] allocator-counter extension unavailable ({exc}); "
                  "using torch.cuda.*memory_stats", file=sys.stderr)
            _alloc = None
    return _alloc


class StepMemoryTracker:
    __slots__ = ("model_id", "device", "is_cuda", "index")

    def __init__(self, model):
        self.model_id = id(model)
        self.is_cuda = False
        self.index = 0
        self.device = None
        if disabled():
            return
        import torch

        try:
            self.device = next(model.parameters()).device
        except StopIteration:
            self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.is_cuda = self.device.type == "cuda"
        if self.is_cuda:
            self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()

    def reset(self) -> None:
        if disabled() or not self.is_cuda:
            return
        ext = _alloc_ext()
        if ext is not None:
            ext.reset_peaks(self.index)
        else:
            import torch

            torch.cuda.reset_peak_memory_stats(self.device)

    def record(self) -> None:
        if disabled():
            return
        if self.is_cuda:
            ext = _alloc_ext()
            if ext is not None:
                a, r = ext.peak_bytes(self.index)
            else:
                import torch

                a = torch.cuda.max_memory_allocated(self.device)
                r = torch.cuda.max_memory_reserved(self.device)
            _pending[self.model_id] = (int(a), int(r), str(self.device))
        else:
            _pending[self.model_id] = (None, None, str(self.device))


def take_pending(model) -> Optional[Tuple[Optional[int], Optional[int], str]]:
    return _pending.pop(id(model), None)


def flush_step_memory_buffer(model, step: int) -> None:
    """Seam compatibility: the peaks ride in the step record that
    ``flush_step_events`` commits (utils/step_memory.py:93-110)."""
    return None


__all__ = ["StepMemoryTracker", "flush_step_memory_buffer", "take_pending"]
