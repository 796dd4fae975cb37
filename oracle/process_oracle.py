"""Oracle: Process reduce + diagnosis.  TEST INFRASTRUCTURE ONLY.

Restates the SQL aggregates of
``src/traceml/reporting/sections/process/loader.py:31-230`` (latest N rows per
global rank; pooled and per-rank AVG / MAX; overhang = MAX(reserved / used)
over rows with used > 0) and the rule engine of
``src/traceml/diagnostics/process/context.py:140-340``, ``rules.py:57-345``,
``api.py:52-118``, ``policy.py:21-31``, ``diagnostics/bands.py:21-34``.

Input: ``rows_by_rank[rank]`` = list (insertion order) of dicts with keys
``ts, cpu, cpu_cores, ram_used, ram_total, gpu_available, gpu_count,
mem_used, mem_reserved, mem_total`` (None allowed for the GPU fields) -- the
numeric columns of ``process_samples``
(``aggregator/sqlite_writers/process.py:169-190``).
"""

from __future__ import annotations

from typing import Any, Dict, List, Optional

ISSUE_PRIORITY = {  # rules.py:297-304
    "VERY_HIGH_PROCESS_GPU_MEMORY": 0, "HIGH_PROCESS_GPU_MEMORY": 1,
    "GPU_MEMORY_RESERVED_OVERHANG": 2, "RANK_GPU_MEMORY_IMBALANCE": 3,
    "HIGH_PROCESS_RSS": 4, "HIGH_PROCESS_CPU": 5,
}


def _avg(vals):
    v = [float(x) for x in vals if x is not None]
    return (sum(v) / len(v)) if v else None


def _max(vals):
    v = [x for x in vals if x is not None]
    return max(v) if v else None


def _f(x):
    return float(x) if x is not None else None


def _aggregate(rows: List[Dict[str, Any]]) -> Dict[str, Any]:
    col = lambda k: [r.get(k) for r in rows]  # noqa: E731
    cores = _max(col("cpu_cores"))
    avail = _max([None if r.get("gpu_available") is None
                  else int(bool(r.get("gpu_available"))) for r in rows])
    count = _max(col("gpu_count"))
    return {
        "cpu_avg_percent": _f(_avg(col("cpu"))),
        "cpu_peak_percent": _f(_max(col("cpu"))),
        "cpu_logical_core_count": int(cores) if cores is not None else None,
        "ram_avg_bytes": _f(_avg(col("ram_used"))),
        "ram_peak_bytes": _f(_max(col("ram_used"))),
        "ram_total_bytes": _f(_max(col("ram_total"))),
        "gpu_available": bool(avail) if avail is not None else None,
        "gpu_count": int(count) if count is not None else None,
        "gpu_mem_used_avg_bytes": _f(_avg(col("mem_used"))),
        "gpu_mem_used_peak_bytes": _f(_max(col("mem_used"))),
        "gpu_mem_reserved_avg_bytes": _f(_avg(col("mem_reserved"))),
        "gpu_mem_reserved_peak_bytes": _f(_max(col("mem_reserved"))),
        "gpu_mem_total_bytes": _f(_max(col("mem_total"))),
    }


def load_section(rows_by_rank, max_rows: int = 10_000):
    """loader.py:56-230."""
    n = max(1, int(max_rows))
    recent = {int(r): list(rows)[-n:] for r, rows in rows_by_rank.items()}
    pooled = [row for r in sorted(recent) for row in recent[r]]
    agg = _aggregate(pooled)
    ts = [row.get("ts") for row in pooled if row.get("ts") is not None]
    agg.update({
        "first_ts": float(min(ts)) if ts else None,
        "last_ts": float(max(ts)) if ts else None,
        "process_samples": len(pooled),
        "distinct_global_ranks": sum(1 for r in recent if recent[r]),
    })
    per_rank = {}
    for r in sorted(recent):
        if not recent[r]:
            continue
        a = _aggregate(recent[r])
        ratios = [row["mem_reserved"] / row["mem_used"] for row in recent[r]
                  if row.get("mem_used") is not None and row["mem_used"] > 0
                  and row.get("mem_reserved") is not None]
        a["gpu_mem_reserved_overhang_ratio"] = float(max(ratios)) if ratios else None
        a["global_rank"] = r
        per_rank[r] = a
    return {"aggregate": agg, "per_global_rank": per_rank}


def classify(value, low_below=None, high_at=None, very_high_at=None):
    """diagnostics/bands.py:21-34."""
    if value is None:
        return None
    v = float(value)
    if very_high_at is not None and v >= very_high_at:
        return "very_high"
    if high_at is not None and v >= high_at:
        return "high"
    if low_below is not None and v < low_below:
        return "low"
    return "normal"


def _frac(num, den):
    if num is None or den is None or float(den) <= 0.0:
        return None
    return max(0.0, float(num) / float(den))


def _best_rank(per_rank, key):  # context.py:151-171 (first strict max)
    best, best_v = None, None
    for r, item in per_rank.items():
        v = item.get(key)
        if v is None:
            continue
        if best_v is None or float(v) > best_v:
            best, best_v = int(r), float(v)
    return best


def _imbalance(per_rank, key):  # context.py:174-194
    vals = [float(i[key]) for i in per_rank.values() if i.get(key) is not None]
    if len(vals) < 2:
        return None
    mx, mn = max(vals), min(vals)
    if mx <= 0.0:
        return 0.0
    return max(0.0, (mx - mn) / mx)


def signals(data) -> Dict[str, Any]:
    """context.py:242-340 (build_process_summary_signals)."""
    agg, per_rank = data["aggregate"], data["per_global_rank"]
    lh_rank, lh_bytes = None, None  # context.py:197-216
    for r, it in per_rank.items():
        tot, resv = it.get("gpu_mem_total_bytes"), it.get("gpu_mem_reserved_peak_bytes")
        if tot is None or resv is None:
            continue
        head = max(float(tot) - float(resv), 0.0)
        if lh_bytes is None or head < lh_bytes:
            lh_rank, lh_bytes = int(r), head
    oh_ratio, oh_rank = None, None  # context.py:219-239
    for r, it in per_rank.items():
        ratio = it.get("gpu_mem_reserved_overhang_ratio")
        if ratio is None:
            ratio = _frac(it.get("gpu_mem_reserved_peak_bytes"),
                          it.get("gpu_mem_used_peak_bytes"))
        if ratio is None:
            continue
        if oh_ratio is None or ratio > oh_ratio:
            oh_ratio, oh_rank = float(ratio), int(r)
    cpu_frac = None
    if (agg["cpu_avg_percent"] is not None and agg["cpu_logical_core_count"] is not None
            and agg["cpu_logical_core_count"] > 0):
        cpu_frac = max(0.0, float(agg["cpu_avg_percent"])
                       / (100.0 * float(agg["cpu_logical_core_count"])))
    ram_frac = _frac(agg["ram_peak_bytes"], agg["ram_total_bytes"])
    used_frac = _frac(agg["gpu_mem_used_peak_bytes"], agg["gpu_mem_total_bytes"])
    resv_frac = _frac(agg["gpu_mem_reserved_peak_bytes"], agg["gpu_mem_total_bytes"])
    used_imb = _imbalance(per_rank, "gpu_mem_used_peak_bytes")
    resv_imb = _imbalance(per_rank, "gpu_mem_reserved_peak_bytes")
    pc = lambda v: v * 100.0 if v is not None else None  # noqa: E731
    first, last = agg["first_ts"], agg["last_ts"]
    duration = None if (first is None or last is None or last < first) else last - first
    return {
        "duration_s": duration, "samples": int(agg["process_samples"]),
        "distinct_ranks": int(agg["distinct_global_ranks"]),
        "cpu_avg_percent": agg["cpu_avg_percent"],
        "cpu_logical_core_count": agg["cpu_logical_core_count"],
        "cpu_capacity_percent": pc(cpu_frac),
        "ram_peak_percent": pc(ram_frac),
        "gpu_mem_used_peak_percent": pc(used_frac),
        "gpu_mem_reserved_peak_percent": pc(resv_frac),
        "gpu_mem_reserved_overhang_ratio": oh_ratio,
        "highest_overhang_rank": oh_rank,
        "highest_rss_rank": _best_rank(per_rank, "ram_peak_bytes"),
        "highest_used_rank": _best_rank(per_rank, "gpu_mem_used_peak_bytes"),
        "highest_reserved_rank": _best_rank(per_rank, "gpu_mem_reserved_peak_bytes"),
        "least_headroom_rank": lh_rank, "least_headroom_bytes": lh_bytes,
        "rank_gpu_used_imbalance_percent": pc(used_imb),
        "rank_gpu_reserved_imbalance_percent": pc(resv_imb),
    }


def _issue(kind, severity, summary, action, metric, phase, score, ranks, evidence):
    return {"kind": kind, "status": kind.replace("_", " "), "severity": severity,
            "summary": summary, "action": action, "metric": metric, "phase": phase,
            "score": float(score) if score is not None else None,
            "share_pct": None, "skew_pct": None,
            "ranks": tuple(int(r) for r in ranks), "evidence": dict(evidence or {})}


def _pct(v):
    return "n/a" if v is None else f"{float(v):.1f}%"


def run_rules(s) -> List[Dict[str, Any]]:
    """rules.py:71-345."""
    out = []
    if s["gpu_mem_reserved_peak_percent"] is not None:  # rules.py:57-68
        pct, metric, rank = (s["gpu_mem_reserved_peak_percent"],
                             "gpu_mem_reserved_peak_percent", s["highest_reserved_rank"])
    else:
        pct, metric, rank = (s["gpu_mem_used_peak_percent"],
                             "gpu_mem_used_peak_percent", s["highest_used_rank"])
    band = classify(pct, low_below=30.0, high_at=80.0, very_high_at=90.0)
    on_rank = "" if rank is None else f" on rank {int(rank)}"
    ev = {"gpu_mem_used_peak_percent": s["gpu_mem_used_peak_percent"],
          "gpu_mem_reserved_peak_percent": s["gpu_mem_reserved_peak_percent"],
          "rank": rank}
    if band == "very_high":
        out.append(_issue(
            "VERY_HIGH_PROCESS_GPU_MEMORY", "crit",
            f"Process GPU memory was very high, peaking at {_pct(pct)}{on_rank}.",
            "Reduce traced process GPU memory pressure.", metric, "gpu_memory",
            pct, () if rank is None else (rank,), ev))
    if band == "high":
        out.append(_issue(
            "HIGH_PROCESS_GPU_MEMORY", "warn",
            f"Process GPU memory was high, peaking at {_pct(pct)}{on_rank}.",
            "Watch traced process GPU memory headroom.", metric, "gpu_memory",
            pct, () if rank is None else (rank,), ev))
    ratio = s["gpu_mem_reserved_overhang_ratio"]
    if classify(ratio, high_at=1.5) == "high":
        r = s["highest_overhang_rank"]
        out.append(_issue(
            "GPU_MEMORY_RESERVED_OVERHANG", "warn",
            f"Reserved GPU memory was {float(ratio):.2f}x active use.",
            "Inspect allocator behavior or retained tensors.",
            "gpu_mem_reserved_peak_bytes", "gpu_memory", ratio,
            () if r is None else (r,),
            {"gpu_mem_reserved_overhang_ratio": float(ratio),
             "highest_overhang_rank": r}))
    ipct, imetric = s["rank_gpu_reserved_imbalance_percent"], "rank_gpu_reserved_imbalance_percent"
    iranks = tuple(r for r in (s["highest_reserved_rank"], s["least_headroom_rank"])
                   if r is not None)
    if ipct is None:
        ipct, imetric = s["rank_gpu_used_imbalance_percent"], "rank_gpu_used_imbalance_percent"
        iranks = tuple(r for r in (s["highest_used_rank"], s["least_headroom_rank"])
                       if r is not None)
    if classify(ipct, high_at=30.0) == "high":
        out.append(_issue(
            "RANK_GPU_MEMORY_IMBALANCE", "warn",
            f"Process GPU memory differed by {_pct(ipct)} across ranks.",
            "Inspect per-rank workload and memory behavior.", imetric,
            "gpu_memory", ipct, iranks,
            {"rank_gpu_used_imbalance_percent": s["rank_gpu_used_imbalance_percent"],
             "rank_gpu_reserved_imbalance_percent": s["rank_gpu_reserved_imbalance_percent"],
             "highest_used_rank": s["highest_used_rank"],
             "highest_reserved_rank": s["highest_reserved_rank"],
             "least_headroom_rank": s["least_headroom_rank"]}))
    rpct = s["ram_peak_percent"]
    if classify(rpct, low_below=30.0, high_at=80.0) == "high":
        r = s["highest_rss_rank"]
        out.append(_issue(
            "HIGH_PROCESS_RSS", "warn",
            f"Process RSS was high, peaking at {_pct(rpct)}.",
            "Reduce traced process host memory pressure.", "ram_peak_percent",
            "ram", rpct, () if r is None else (r,),
            {"ram_peak_percent": rpct, "highest_rss_rank": r}))
    cpct = s["cpu_capacity_percent"]
    if classify(cpct, low_below=30.0, high_at=80.0) == "high":
        out.append(_issue(
            "HIGH_PROCESS_CPU", "warn",
            f"Process CPU averaged {_pct(cpct)} of capacity.",
            "Inspect data loading, preprocessing, or host contention.",
            "cpu_capacity_percent", "cpu", cpct, (),
            {"cpu_avg_percent": s["cpu_avg_percent"],
             "cpu_logical_core_count": s["cpu_logical_core_count"],
             "cpu_capacity_percent": cpct}))
    return sorted(out, key=lambda i: (ISSUE_PRIORITY.get(i["kind"], 999),
                                      -(float(i["score"] or 0.0))))


def diagnose(data) -> Dict[str, Any]:
    """api.py:52-118."""
    s = signals(data)
    issues = run_rules(s) if s["samples"] > 0 else []
    if issues:
        top = issues[0]
        primary = {"kind": top["kind"], "severity": top["severity"],
                   "status": top["status"], "reason": top["summary"],
                   "action": top["action"], "samples_used": s["samples"]}
    elif s["samples"] <= 0:
        primary = {"kind": "NO_DATA", "severity": "info", "status": "NO DATA",
                   "reason": "No traced process telemetry was recorded.",
                   "action": "Collect process telemetry for workload-local context.",
                   "samples_used": s["samples"]}
    else:
        has_gpu = (s["gpu_mem_used_peak_percent"] is not None
                   or s["gpu_mem_reserved_peak_percent"] is not None)
        primary = {"kind": "NORMAL", "severity": "info", "status": "NORMAL",
                   "reason": ("Process CPU, RSS, and GPU memory showed no pressure."
                              if has_gpu else "Process CPU and RSS showed no pressure."),
                   "action": "Use training diagnostics for model-level bottlenecks.",
                   "samples_used": s["samples"]}
    return {"primary": primary, "issues": tuple(issues), "signals": s}


def process_section(rows_by_rank, *, max_rows=10_000):
    data = load_section(rows_by_rank, max_rows)
    return {"data": data, "diagnosis": diagnose(data)}
