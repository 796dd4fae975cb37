"""Binary record layouts shared by the CUDA kernels, the C-ABI and the host.

These are the layouts fixed by SURVEY.md section 8(d); ``include/traceml_b200.h``
declares the same structs for C callers and ``csrc/tml_records.cuh`` for the
kernels.  The numpy dtypes here are bit-for-bit views of those structs.

StepRecord (128 B / step / rank) replaces the reference's ``StepTimeBatch`` of
``TimeEvent`` objects plus its ``StepMemoryEvent``
(``src/traceml/utils/timing.py:44-105``, ``src/traceml/utils/step_memory.py:17-27``).
ProcRecord (64 B / sample / rank) replaces ``ProcessSample``
(``src/traceml/samplers/schema/process.py:89-150``).
"""

from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np

# phase ids == column order of WindowRow
PHASE_DATALOADER = 0
PHASE_H2D = 1
PHASE_FORWARD = 2
PHASE_BACKWARD = 3
PHASE_OPTIMIZER = 4
PHASE_STEP = 5
N_PHASES = 6

# canonical event names, src/traceml/renderers/step_time/compute.py:24-31
PHASE_EVENT_NAMES = (
    "_traceml_internal:dataloader_next",
    "_traceml_internal:h2d_time",
    "_traceml_internal:forward_time",
    "_traceml_internal:backward_time",
    "_traceml_internal:optimizer_step",
    "_traceml_internal:step_time",
)

FLAG_HAS_MEM = 1  # peak_alloc / peak_resv are real (device was CUDA)

STEP_RECORD_DTYPE = np.dtype(
    [
        ("step", "<u8"),
        ("dur_ns", "<u8", (N_PHASES,)),
        ("n_calls", "<u4", (N_PHASES,)),
        ("peak_alloc", "<u8"),
        ("peak_resv", "<u8"),
        ("host_ts", "<f8"),
        ("gpu_mask", "<u4"),
        ("flags", "<u4"),
        ("seq", "<u8"),
        ("_pad", "<u8"),
    ]
)
assert STEP_RECORD_DTYPE.itemsize == 128

WINDOW_ROW_DTYPE = np.dtype(
    [
        ("dataloader", "<f8"),
        ("h2d", "<f8"),
        ("forward", "<f8"),
        ("backward", "<f8"),
        ("optimizer", "<f8"),
        ("step_wall", "<f8"),
        ("peak_alloc", "<f8"),
        ("peak_resv", "<f8"),
    ]
)
assert WINDOW_ROW_DTYPE.itemsize == 64

PROC_FLAG_GPU_AVAILABLE = 1
PROC_FLAG_HAS_GPU_METRICS = 2

PROC_RECORD_DTYPE = np.dtype(
    [
        ("seq", "<u8"),
        ("ts", "<f8"),
        ("cpu_pct", "<f8"),
        ("rss", "<u8"),
        ("mem_alloc", "<u8"),
        ("mem_resv", "<u8"),
        ("mem_total", "<u8"),
        ("flags", "<u4"),
        ("cpu_cores", "<u4"),
    ]
)
assert PROC_RECORD_DTYPE.itemsize == 64


def ns_to_ms(ns) -> float:
    """The one ns -> ms conversion used everywhere (a true division so the
    result is the correctly rounded quotient, same as the kernels' ``__ddiv_rn``)."""
    return float(ns) / 1.0e6


def step_record_to_wire(rec, *, seq: Optional[int] = None, device: str = "cuda:0") -> Dict[str, Any]:
    """One StepRecord -> the reference's step-time wire row
    (``StepTimeEventSample.to_wire``, samplers/schema/step_time_schema.py:85-96)."""
    events: Dict[str, Dict[str, Dict[str, Any]]] = {}
    mask = int(rec["gpu_mask"])
    for p in range(N_PHASES):
        calls = int(rec["n_calls"][p])
        if calls <= 0:
            continue
        is_gpu = bool((mask >> p) & 1)
        events[PHASE_EVENT_NAMES[p]] = {
            (device if is_gpu else "cpu"): {
                "is_gpu": is_gpu,
                "duration_ms": ns_to_ms(int(rec["dur_ns"][p])),
                "n_calls": calls,
            }
        }
    return {
        "seq": int(rec["seq"] if seq is None else seq),
        "timestamp": float(rec["host_ts"]),
        "step": int(rec["step"]),
        "events": events,
    }


def step_record_to_memory_wire(rec, *, seq: Optional[int] = None, device: str = "cuda:0",
                               model_id: int = 0, ts: Optional[float] = None) -> Dict[str, Any]:
    """One StepRecord -> the step-memory wire row
    (``StepMemorySample.to_wire``, samplers/schema/step_memory.py:41-58)."""
    has = bool(int(rec["flags"]) & FLAG_HAS_MEM)
    return {
        "seq": int(rec["seq"] if seq is None else seq),
        "ts": float(rec["host_ts"] if ts is None else ts),
        "model_id": int(model_id),
        "device": device if has else "cpu",
        "step": int(rec["step"]),
        "peak_alloc": float(int(rec["peak_alloc"])) if has else None,
        "peak_resv": float(int(rec["peak_resv"])) if has else None,
    }


def proc_record_to_wire(rec, *, pid: int = 0, ram_total: float = 0.0, gpu_count: int = 0,
                        device_index: int = 0) -> Dict[str, Any]:
    """One ProcRecord -> the process wire row
    (``ProcessSample.to_wire``, samplers/schema/process.py:139-150)."""
    flags = int(rec["flags"])
    gpu = None
    if flags & PROC_FLAG_HAS_GPU_METRICS:
        gpu = {
            "device": int(device_index),
            "mem_used": float(int(rec["mem_alloc"])),
            "mem_reserved": float(int(rec["mem_resv"])),
            "mem_total": float(int(rec["mem_total"])),
        }
    avail = bool(flags & PROC_FLAG_GPU_AVAILABLE)
    return {
        "seq": int(rec["seq"]),
        "ts": float(rec["ts"]),
        "pid": int(pid),
        "cpu": float(rec["cpu_pct"]),
        "cpu_cores": int(rec["cpu_cores"]),
        "ram_used": float(int(rec["rss"])),
        "ram_total": float(ram_total),
        "gpu_available": avail,
        "gpu_count": int(gpu_count) if avail else 0,
        "gpu": gpu,
    }


def records_to_time_rows(records: np.ndarray, max_rows: int) -> List[Dict[str, Any]]:
    """The rows the reference loader would hand to its reduce for one rank:
    ``ORDER BY step DESC, id DESC LIMIT max_rows`` with the JSON already parsed
    (reporting/sections/step_time/loader.py:44-72)."""
    n = len(records)
    order = sorted(range(n), key=lambda i: (int(records["step"][i]), i), reverse=True)
    rows = []
    for i in order[: max(1, int(max_rows))]:
        w = step_record_to_wire(records[i])
        rows.append({"step": w["step"], "events": w["events"]})
    return rows


__all__ = [
    "STEP_RECORD_DTYPE", "WINDOW_ROW_DTYPE", "PROC_RECORD_DTYPE", "N_PHASES",
    "PHASE_EVENT_NAMES", "PHASE_DATALOADER", "PHASE_H2D", "PHASE_FORWARD",
    "PHASE_BACKWARD", "PHASE_OPTIMIZER", "PHASE_STEP", "FLAG_HAS_MEM",
    "PROC_FLAG_GPU_AVAILABLE", "PROC_FLAG_HAS_GPU_METRICS", "ns_to_ms",
    "step_record_to_wire", "step_record_to_memory_wire", "proc_record_to_wire",
    "records_to_time_rows",
]
