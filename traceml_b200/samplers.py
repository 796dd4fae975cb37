"""Sampler-side views: device records -> the reference's wire rows.

Replaces ``StepTimeSampler`` / ``StepMemorySampler`` / ``ProcessSampler``
(``src/traceml/samplers/{step_time,step_memory,process}_sampler.py``).  The
step samplers become a drain of the host-mapped record mirror (no CUDA call,
no event queries); the process sampler still reads CPU% / RSS on the host
(psutil) but commits the sample through ``tml_proc_commit`` into the device
proc ring so the Process reduce runs on the GPU with the rest.
"""
from __future__ import annotations

import os
import time
from typing import Any, Dict, List, Optional

from .records import (PROC_FLAG_GPU_AVAILABLE, PROC_FLAG_HAS_GPU_METRICS, proc_record_to_wire,
                      step_record_to_memory_wire, step_record_to_wire)


def drain_to_wire(engine, device: Optional[str] = None) -> Dict[str, Any]:
    """Completed records since the last call as wire rows
    (samplers/schema/step_time_schema.py:85-96, step_memory.py:41-58, process.py:139-150)."""
    dev = device or f"cuda:{engine.device}"
    recs, dropped = engine.drain()
    procs, pdropped = engine.proc_drain()
    now = time.time()
    return {
        "step_time": [step_record_to_wire(r, device=dev) for r in recs],
        "step_memory": [step_record_to_memory_wire(r, device=dev, ts=now) for r in recs],
        "process": [proc_record_to_wire(p, pid=os.getpid(), device_index=engine.device) for p in procs],
        "dropped": int(dropped) + int(pdropped),
    }


class ProcessProbe:
    """One process sample per call (samplers/process_sampler.py:178-238)."""

    def __init__(self):
        import psutil

        self.proc = psutil.Process(os.getpid())
        self.proc.cpu_percent(interval=None)  # warm-up, as the reference does
        self.cores = psutil.cpu_count(logical=True) or 0
        self.ram_total = float(psutil.virtual_memory().total)
        self.seq = 0
        self._stream = None
        self._total = None

    def _cuda_safe(self) -> bool:
        # never touch CUDA before dist.init_process_group() in a distributed job
        # (process_sampler.py:150-158)
        if int(os.environ.get("WORLD_SIZE", "1") or 1) <= 1:
            return True
        try:
            import torch.distributed as dist

            return dist.is_available() and dist.is_initialized()
        except Exception:
            return False

    def sample(self, engine) -> None:
        import torch

        from .utils.step_memory import _alloc_ext

        self.seq += 1
        cpu = float(self.proc.cpu_percent(interval=None))
        rss = int(self.proc.memory_info().rss)
        flags, used, resv, total = 0, 0, 0, 0
        if self._cuda_safe() and torch.cuda.is_available():
            flags |= PROC_FLAG_GPU_AVAILABLE | PROC_FLAG_HAS_GPU_METRICS
            ext = _alloc_ext()
            if ext is not None:
                used, resv = ext.current_bytes(engine.device)
            else:
                used = torch.cuda.memory_allocated(engine.device)
                resv = torch.cuda.memory_reserved(engine.device)
            if self._total is None:
                self._total = int(torch.cuda.get_device_properties(engine.device).total_memory)
            total = self._total
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=engine.device)  # side stream, never the training stream
        engine.proc_commit(self.seq, time.time(), cpu, rss, int(used), int(resv), int(total), flags,
                           int(self.cores), int(self._stream.cuda_stream))


__all__ = ["drain_to_wire", "ProcessProbe"]
