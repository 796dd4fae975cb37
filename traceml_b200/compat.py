"""SQLite compatibility writer (SURVEY section 8(f) "next" row 1).

Writes the reference aggregator's projection tables --
``step_time_samples`` (``events_json`` TEXT), ``step_memory_samples``,
``process_samples`` from drained device records, ``system_samples`` /
``system_gpu_samples`` from the host snapshot -- so the *kept* consumers
(live CLI / dashboard computers, ``traceml compare`` / ``inspect``, and the
reference's own final-report sections) keep working on a database this engine
produced.  Format only, no arithmetic; the schemas follow
``aggregator/sqlite_writers/step_time.py:159-247``, ``step_memory.py:133-219``,
``process.py:157-242`` and the row shaping ``build_rows`` of each.

Use as a sink of :class:`traceml_b200.runtime.TraceMLRuntime`::

    writer = SQLiteCompatWriter(path, identity)
    TraceMLRuntime(sinks=[writer]).start()
"""
from __future__ import annotations

import json
import sqlite3
import time
from typing import Any, Dict, Iterable, List, Optional

_ID_COLS = ("rank", "global_rank", "local_rank", "world_size", "local_world_size", "node_rank", "hostname")

_SCHEMA = (
    """CREATE TABLE IF NOT EXISTS step_time_samples (
        id INTEGER PRIMARY KEY AUTOINCREMENT, recv_ts_ns INTEGER NOT NULL, rank INTEGER,
        global_rank INTEGER, local_rank INTEGER, world_size INTEGER, local_world_size INTEGER,
        node_rank INTEGER, hostname TEXT, runtime_pid INTEGER, sample_ts_s REAL, seq INTEGER,
        step INTEGER, events_json TEXT NOT NULL);""",
    """CREATE TABLE IF NOT EXISTS step_memory_samples (
        id INTEGER PRIMARY KEY AUTOINCREMENT, recv_ts_ns INTEGER NOT NULL, rank INTEGER,
        global_rank INTEGER, local_rank INTEGER, world_size INTEGER, local_world_size INTEGER,
        node_rank INTEGER, hostname TEXT, sample_ts_s REAL, seq INTEGER, model_id INTEGER,
        device TEXT, step INTEGER, peak_alloc_bytes REAL, peak_reserved_bytes REAL);""",
    """CREATE TABLE IF NOT EXISTS process_samples (
        id INTEGER PRIMARY KEY AUTOINCREMENT, recv_ts_ns INTEGER NOT NULL, rank INTEGER,
        global_rank INTEGER, local_rank INTEGER, world_size INTEGER, local_world_size INTEGER,
        node_rank INTEGER, hostname TEXT, sample_ts_s REAL, seq INTEGER, cpu_percent REAL,
        cpu_logical_core_count INTEGER, ram_used_bytes REAL, ram_total_bytes REAL,
        gpu_available INTEGER, gpu_count INTEGER, gpu_device_index INTEGER,
        gpu_mem_used_bytes REAL, gpu_mem_reserved_bytes REAL, gpu_mem_total_bytes REAL);""",
    """CREATE TABLE IF NOT EXISTS system_samples (
        id INTEGER PRIMARY KEY AUTOINCREMENT, recv_ts_ns INTEGER NOT NULL, global_rank INTEGER,
        local_rank INTEGER, world_size INTEGER, local_world_size INTEGER, node_rank INTEGER, hostname TEXT,
        sample_ts_s REAL, seq INTEGER, cpu_percent REAL, ram_used_bytes REAL, ram_total_bytes REAL,
        gpu_available INTEGER, gpu_count INTEGER, gpu_util_avg REAL, gpu_util_peak REAL,
        gpu_mem_used_avg_bytes REAL, gpu_mem_used_peak_bytes REAL, gpu_temp_avg_c REAL, gpu_temp_peak_c REAL,
        gpu_power_avg_w REAL, gpu_power_peak_w REAL);""",
    """CREATE TABLE IF NOT EXISTS system_gpu_samples (
        id INTEGER PRIMARY KEY AUTOINCREMENT, recv_ts_ns INTEGER NOT NULL, global_rank INTEGER,
        local_rank INTEGER, world_size INTEGER, local_world_size INTEGER, node_rank INTEGER, hostname TEXT,
        sample_ts_s REAL, seq INTEGER, gpu_idx INTEGER NOT NULL, util REAL, mem_used_bytes REAL,
        mem_total_bytes REAL, temperature_c REAL, power_usage_w REAL, power_limit_w REAL);""",
    "CREATE INDEX IF NOT EXISTS idx_system_samples_node_ts ON system_samples(node_rank, sample_ts_s, id);",
    "CREATE INDEX IF NOT EXISTS idx_system_gpu_samples_global_gpu_ts ON system_gpu_samples(global_rank, gpu_idx, sample_ts_s, id);",
    "CREATE INDEX IF NOT EXISTS idx_step_time_samples_rank_step_ts ON step_time_samples(rank, step, sample_ts_s, id);",
    "CREATE INDEX IF NOT EXISTS idx_step_time_samples_global_rank_step_ts ON step_time_samples(global_rank, step, sample_ts_s, id);",
    "CREATE INDEX IF NOT EXISTS idx_step_time_samples_step_rank ON step_time_samples(step, rank, id);",
    "CREATE INDEX IF NOT EXISTS idx_step_memory_samples_rank_step_ts ON step_memory_samples(rank, step, sample_ts_s, id);",
    "CREATE INDEX IF NOT EXISTS idx_step_memory_samples_global_rank_step_ts ON step_memory_samples(global_rank, step, sample_ts_s, id);",
    "CREATE INDEX IF NOT EXISTS idx_process_samples_rank_ts ON process_samples(rank, sample_ts_s, id);",
    "CREATE INDEX IF NOT EXISTS idx_process_samples_global_rank_ts ON process_samples(global_rank, sample_ts_s, id);",
)


class SQLiteCompatWriter:
    def __init__(self, db_path: str, identity: Dict[str, Any], pid: Optional[int] = None):
        self.conn = sqlite3.connect(db_path, check_same_thread=False)
        for stmt in _SCHEMA:
            self.conn.execute(stmt)
        self.conn.commit()
        g = identity.get("global_rank")
        self._ident = (g, g, identity.get("local_rank"), identity.get("world_size"),
                       identity.get("local_world_size"), identity.get("node_rank"),
                       identity.get("hostname"))
        self.pid = pid

    # sink protocol of TraceMLRuntime
    def __call__(self, kind: str, rows: Iterable[Dict[str, Any]]) -> None:
        getattr(self, f"write_{kind}")(list(rows))

    def write_step_time(self, rows: List[Dict[str, Any]]) -> None:
        now = time.time_ns()
        self.conn.executemany(
            "INSERT INTO step_time_samples(recv_ts_ns, rank, global_rank, local_rank, world_size, "
            "local_world_size, node_rank, hostname, runtime_pid, sample_ts_s, seq, step, events_json) "
            "VALUES (?,?,?,?,?,?,?,?,?,?,?,?,?);",
            [(now, *self._ident, self.pid, float(r["timestamp"]), int(r["seq"]), int(r["step"]),
              json.dumps(r["events"], separators=(",", ":"), sort_keys=True)) for r in rows])
        self.conn.commit()

    def write_step_memory(self, rows: List[Dict[str, Any]]) -> None:
        now = time.time_ns()
        self.conn.executemany(
            "INSERT INTO step_memory_samples(recv_ts_ns, rank, global_rank, local_rank, world_size, "
            "local_world_size, node_rank, hostname, sample_ts_s, seq, model_id, device, step, "
            "peak_alloc_bytes, peak_reserved_bytes) VALUES (?,?,?,?,?,?,?,?,?,?,?,?,?,?,?);",
            [(now, *self._ident, float(r["ts"]), int(r["seq"]), r.get("model_id"), r.get("device"),
              int(r["step"]), r.get("peak_alloc"), r.get("peak_resv")) for r in rows])
        self.conn.commit()

    def write_process(self, rows: List[Dict[str, Any]]) -> None:
        now = time.time_ns()
        out = []
        for r in rows:
            g = r.get("gpu") or {}
            avail = r.get("gpu_available")
            out.append((now, *self._ident, float(r["ts"]), int(r["seq"]), float(r["cpu"]),
                        int(r["cpu_cores"]), float(r["ram_used"]), float(r["ram_total"]),
                        (1 if avail is True else 0 if avail is False else None), int(r["gpu_count"]),
                        g.get("device"), g.get("mem_used"), g.get("mem_reserved"), g.get("mem_total")))
        self.conn.executemany(
            "INSERT INTO process_samples(recv_ts_ns, rank, global_rank, local_rank, world_size, "
            "local_world_size, node_rank, hostname, sample_ts_s, seq, cpu_percent, "
            "cpu_logical_core_count, ram_used_bytes, ram_total_bytes, gpu_available, gpu_count, "
            "gpu_device_index, gpu_mem_used_bytes, gpu_mem_reserved_bytes, gpu_mem_total_bytes) "
            "VALUES (?,?,?,?,?,?,?,?,?,?,?,?,?,?,?,?,?,?,?,?);", out)
        self.conn.commit()

    def write_system(self, rows: List[Dict[str, Any]]) -> None:
        """aggregator/sqlite_writers/system.py:280-480: one host row + one row per GPU; the legacy
        ``rank`` column is not stored for system tables."""
        now = time.time_ns()
        ident = self._ident[1:]
        host, gpus = [], []
        for r in rows:
            ts, seq = float(r["ts"]), int(r["seq"])
            cols = list(zip(*[g for g in r.get("gpus") or [] if isinstance(g, list) and len(g) >= 6])) or [[]] * 6
            for i, g in enumerate(r.get("gpus") or []):
                if isinstance(g, list) and len(g) >= 6:
                    gpus.append((now, *ident, ts, seq, i, *[float(v) for v in g[:6]]))
            avg = lambda v: (sum(v) / len(v)) if v else None  # noqa: E731
            peak = lambda v: max(v) if v else None            # noqa: E731
            util, mem, temp, power = list(cols[0]), list(cols[1]), list(cols[3]), list(cols[4])
            avail = r.get("gpu_available")
            host.append((now, *ident, ts, seq, float(r["cpu"]), float(r["ram_used"]), float(r["ram_total"]),
                         (1 if avail is True else 0 if avail is False else None), int(r["gpu_count"]),
                         avg(util), peak(util), avg(mem), peak(mem), avg(temp), peak(temp), avg(power), peak(power)))
        self.conn.executemany(
            "INSERT INTO system_samples(recv_ts_ns, global_rank, local_rank, world_size, local_world_size, node_rank, "
            "hostname, sample_ts_s, seq, cpu_percent, ram_used_bytes, ram_total_bytes, gpu_available, gpu_count, "
            "gpu_util_avg, gpu_util_peak, gpu_mem_used_avg_bytes, gpu_mem_used_peak_bytes, gpu_temp_avg_c, "
            "gpu_temp_peak_c, gpu_power_avg_w, gpu_power_peak_w) VALUES (?,?,?,?,?,?,?,?,?,?,?,?,?,?,?,?,?,?,?,?,?,?);", host)
        self.conn.executemany(
            "INSERT INTO system_gpu_samples(recv_ts_ns, global_rank, local_rank, world_size, local_world_size, node_rank, "
            "hostname, sample_ts_s, seq, gpu_idx, util, mem_used_bytes, mem_total_bytes, temperature_c, power_usage_w, "
            "power_limit_w) VALUES (?,?,?,?,?,?,?,?,?,?,?,?,?,?,?,?);", gpus)
        self.conn.commit()

    def close(self) -> None:
        self.conn.close()


__all__ = ["SQLiteCompatWriter"]
