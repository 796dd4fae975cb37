"""Oracle: the reference's per-step timer / memory / sampler path.  TEST INFRASTRUCTURE ONLY.

A CPU-side restatement of how the reference measures one step, used (a) as the
semantic checker for the host layer (accumulate-by-name, step numbering, wire
rows) and (b) as the ``cpu_baseline`` / ``--impl reference`` arm of bench.py
for the per-step overhead figure.  It uses exactly the mechanisms the
reference uses -- ``time.time()`` pairs, pooled ``torch.cuda.Event`` pairs
recorded on the current stream, ``reset_peak_memory_stats`` /
``max_memory_allocated`` / ``max_memory_reserved``, a bounded ``queue.Queue``
and a sampler that resolves events with ``query()`` -- with none of ours.

Follows (paths under src/traceml/):
  utils/timing.py:44-256            TimeEvent / try_resolve / timed_region / flush
  utils/cuda_event_pool.py:15-71    event free-list (cap 2000, lock per op)
  utils/step_memory.py:30-110       StepMemoryTracker
  sdk/instrumentation.py:160-200    trace_step
  runtime/state.py:54-68            advance_step under an RLock
  samplers/step_time_sampler.py:55-128   drain -> resolve (head of line) -> aggregate
  samplers/step_memory_sampler.py:12-65

Parity status: the reference's own tests never exercise these functions
(SURVEY section 4), so this half of the oracle is pinned only against the
reference code run side by side in tests/golden/make_timer_golden.py on CPU
(host-clock phases); the CUDA-event half is "parity unpinned" by golden
vectors and is compared on the GPU with the physical tolerance of SURVEY 8(d).
"""

from __future__ import annotations

import threading
import time
from collections import defaultdict, deque
from contextlib import contextmanager
from queue import Empty, Full, Queue
from typing import Any, Deque, Dict, List, Optional, Tuple

import torch


class EventPool:
    """cuda_event_pool.py:15-71."""

    def __init__(self, max_size: int = 2000):
        self._pool: Deque[torch.cuda.Event] = deque()
        self._lock = threading.Lock()
        self._max = max_size

    def acquire(self) -> torch.cuda.Event:
        with self._lock:
            if self._pool:
                return self._pool.popleft()
        return torch.cuda.Event(enable_timing=True)

    def release(self, evt: torch.cuda.Event) -> None:
        with self._lock:
            if len(self._pool) < self._max:
                self._pool.append(evt)


class TimeEvent:
    """timing.py:44-90 (a plain class standing in for the dataclass)."""

    __slots__ = ("name", "device", "cpu_start", "cpu_end", "gpu_start", "gpu_end",
                 "gpu_time_ms", "resolved", "step")

    def __init__(self, name, device, cpu_start, cpu_end, gpu_start=None, gpu_end=None):
        self.name, self.device = name, device
        self.cpu_start, self.cpu_end = cpu_start, cpu_end
        self.gpu_start, self.gpu_end = gpu_start, gpu_end
        self.gpu_time_ms = None
        self.resolved = False
        self.step = -1

    def try_resolve(self, pool: EventPool) -> bool:
        if self.resolved:
            return True
        if self.gpu_start and self.gpu_end:
            if self.gpu_end.query():
                self.gpu_time_ms = self.gpu_start.elapsed_time(self.gpu_end)
                pool.release(self.gpu_start)
                pool.release(self.gpu_end)
                self.gpu_start = self.gpu_end = None
                self.resolved = True
        else:
            self.resolved = True
        return self.resolved


class ReferenceTimerPath:
    """One process's worth of the reference's timing state."""

    def __init__(self):
        self.pool = EventPool()
        self.step_queue: Queue = Queue(maxsize=2048)   # timing.py:109
        self.mem_queue: Queue = Queue(maxsize=2048)    # step_memory.py:9
        self.buffer: Deque[TimeEvent] = deque()         # timing.py:111
        self.mem_buffer: Dict[int, Any] = {}
        self.step = 0
        self._lock = threading.RLock()
        self._pending: Deque[Tuple[int, List[TimeEvent]]] = deque()
        self.cuda = torch.cuda.is_available()
        self.dropped = 0

    # ---- timing.py:184-256
    @contextmanager
    def timed_region(self, name: str, scope: str = "step", use_gpu: bool = True):
        cpu_start = time.time()
        if use_gpu and self.cuda:
            device = f"cuda:{torch.cuda.current_device()}"
            start_evt, end_evt = self.pool.acquire(), self.pool.acquire()
            start_evt.record()
        else:
            device, start_evt, end_evt = "cpu", None, None
        try:
            yield
        finally:
            cpu_end = time.time()
            if start_evt is not None:
                end_evt.record()
            evt = TimeEvent(name, device, cpu_start, cpu_end, start_evt, end_evt)
            if scope == "step":
                self.buffer.append(evt)

    # ---- sdk/instrumentation.py:160-200 (+ step_memory.py, flush_buffers.py)
    @contextmanager
    def trace_step(self, model):
        try:
            device = next(model.parameters()).device
        except StopIteration:
            device = torch.device("cuda" if self.cuda else "cpu")
        if device.type == "cuda":
            torch.cuda.reset_peak_memory_stats(device)
        completed = False
        try:
            with self.timed_region("_traceml_internal:step_time", "step", use_gpu=False):
                yield
                completed = True
        finally:
            if completed:
                with self._lock:
                    self.step += 1
            if device.type == "cuda":
                pa = float(torch.cuda.max_memory_allocated(device))
                pr = float(torch.cuda.max_memory_reserved(device))
            else:
                pa = pr = None
            self.mem_buffer[id(model)] = (pa, pr, str(device))
            self.flush(model, self.step)

    def flush(self, model, step: int) -> None:
        mem = self.mem_buffer.pop(id(model), None)
        if mem is not None:
            try:
                self.mem_queue.put_nowait((step, mem))
            except Full:
                self.dropped += 1
        if self.buffer:
            events = []
            while self.buffer:
                e = self.buffer.popleft()
                e.step = step
                events.append(e)
            try:
                self.step_queue.put_nowait((step, events))
            except Full:
                self.dropped += 1

    # ---- samplers/step_time_sampler.py:55-128 + step_memory_sampler.py
    def sample(self) -> Dict[str, List[Dict[str, Any]]]:
        out = {"step_time": [], "step_memory": []}
        while True:
            try:
                self._pending.append(self.step_queue.get_nowait())
            except Empty:
                break
        while self._pending:
            step, events = self._pending[0]
            if not all(e.try_resolve(self.pool) for e in events):
                break  # head-of-line: a step is emitted only when fully resolved
            self._pending.popleft()
            sums: Dict[Tuple[str, str, bool], float] = defaultdict(float)
            calls: Dict[Tuple[str, str, bool], int] = defaultdict(int)
            ts = 0.0
            for e in events:
                ts = float(max(ts, float(e.cpu_end)))
                is_gpu = e.gpu_time_ms is not None
                dur = float(e.gpu_time_ms) if is_gpu else (e.cpu_end - e.cpu_start) * 1000.0
                key = (str(e.name), str(e.device), bool(is_gpu))
                sums[key] += dur
                calls[key] += 1
            ev: Dict[str, Dict[str, Dict[str, Any]]] = defaultdict(dict)
            for (name, device, is_gpu), total in sums.items():
                ev[name][device] = {"is_gpu": bool(is_gpu), "duration_ms": float(total),
                                    "n_calls": int(calls[(name, device, is_gpu)])}
            out["step_time"].append({"seq": 0, "timestamp": ts, "step": int(step),
                                     "events": dict(ev)})
        while True:
            try:
                step, (pa, pr, dev) = self.mem_queue.get_nowait()
            except Empty:
                break
            out["step_memory"].append({"seq": 0, "ts": time.time(), "model_id": 0, "device": dev,
                                       "step": step, "peak_alloc": pa, "peak_resv": pr})
        return out
