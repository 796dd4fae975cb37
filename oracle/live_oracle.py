"""Oracle: the LIVE twin of the Step-Time reduce.  TEST INFRASTRUCTURE ONLY.

Restates ``StepCombinedComputer._compute_impl`` and its helpers
(``src/traceml/renderers/step_time/compute.py:129-315, 432-660``): the
render-tick computation behind the live CLI / dashboard.  It differs from the
final-summary reduce in four ways that matter for parity:

  * candidates are the last ``max(window * lookback_factor, window)`` rows per rank
    (``:366-368``), any row counts (no "usable" filter), first row wins per step id;
  * the window is the last ``window`` step ids present on every rank, walking down
    from ``completed_step = min over ranks of max step`` (``:147, 452-470``);
  * per-rank values are window SUMS of the raw events, ``h2d`` included, and
    ``wait = max(0, step - h2d - fwd - bwd - opt)`` on those sums (``:177-201``);
  * series carry median / worst / sum per step (``:573-596``).

Input: ``rows_by_rank[rank]`` = wire rows ``{"step", "events"}`` in insertion order.
"""

from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import numpy as np

METRIC_KEYS = ("dataloader_fetch", "h2d", "forward", "backward", "optimizer_step", "step_time")
HEATMAP_KEYS = ("dataloader_fetch", "h2d", "forward", "backward", "optimizer_step", "wait_proxy", "step_time")
ALIASES = {
    "dataloader_fetch": "_traceml_internal:dataloader_next",
    "h2d": "_traceml_internal:h2d_time",
    "forward": "_traceml_internal:forward_time",
    "backward": "_traceml_internal:backward_time",
    "optimizer_step": "_traceml_internal:optimizer_step",
    "step_time": "_traceml_internal:step_time",
}


def _sf(x: Any) -> float:
    try:
        v = float(x)
        return v if np.isfinite(v) else 0.0
    except Exception:
        return 0.0


def event_total_ms(payload: Dict[str, Any], key: str) -> float:
    """compute.py:479-500."""
    if not isinstance(payload, dict):
        return 0.0
    p = payload.get(ALIASES.get(key, key))
    if not isinstance(p, dict):
        return 0.0
    return float(sum(_sf(rec.get("duration_ms", 0.0)) for rec in p.values() if isinstance(rec, dict)))


def load_last_steps(rows_by_rank, window: int, lookback_factor: int = 4):
    """compute.py:352-416: last `lookback` rows by (step, id) DESC, replayed ascending,
    first row of a step id wins."""
    lookback = max(window * lookback_factor, window)
    out = {}
    for rank in sorted(rows_by_rank):
        rows = list(rows_by_rank[rank])
        order = sorted(range(len(rows)), key=lambda i: (rows[i]["step"], i), reverse=True)[:lookback]
        order.sort(key=lambda i: rows[i]["step"])  # stable: within a step id, newest first
        m: Dict[int, Dict[str, Any]] = {}
        for i in order:
            s = int(rows[i]["step"])
            if s in m:
                continue
            m[s] = rows[i]["events"]
        if m:
            out[int(rank)] = m
    return out


def common_steps(per_rank_steps, completed_step: int, window: int) -> List[int]:
    """compute.py:452-470."""
    maps = list(per_rank_steps.values())
    if not maps:
        return []
    out: List[int] = []
    for s in range(int(completed_step), -1, -1):
        if all(s in m for m in maps):
            out.append(s)
            if len(out) >= int(window):
                break
    out.reverse()
    return out


def make_metric(key, rank_sums, ranks, coverage, include_series, per_rank_steps, steps):
    """compute.py:533-626."""
    if not ranks:
        return None
    arr = np.array([_sf(rank_sums.get(r, 0.0)) for r in ranks], dtype=np.float64)
    median_total = float(np.median(arr))
    wi = int(np.argmax(arr))
    worst_total, worst_rank = float(arr[wi]), int(ranks[wi])
    if coverage["ranks_present"] <= 1:
        median_total, skew_ratio, skew_pct = worst_total, 0.0, 0.0
    elif median_total > 0.0:
        skew_ratio = worst_total / median_total
        skew_pct = (worst_total - median_total) / median_total
    else:
        skew_ratio = skew_pct = 0.0
    series = None
    if include_series and key != "wait_proxy":
        med, worst, tot = [], [], []
        for st in steps:
            vals = np.array([event_total_ms(per_rank_steps[r].get(int(st), {}), key) for r in ranks],
                            dtype=np.float64)
            med.append(float(np.median(vals)) if vals.size else 0.0)
            worst.append(float(np.max(vals)) if vals.size else 0.0)
            tot.append(float(np.sum(vals)) if vals.size else 0.0)
        series = {"steps": list(steps), "median": med, "worst": worst, "sum": tot}
    return {"metric": str(key), "clock": "mixed", "series": series,
            "summary": {"window_size": int(coverage["expected_steps"]),
                        "steps_used": int(coverage["steps_used"]),
                        "median_total": float(median_total), "worst_total": float(worst_total),
                        "worst_rank": int(worst_rank), "skew_ratio": float(skew_ratio),
                        "skew_pct": float(skew_pct)},
            "coverage": coverage}


def live_step_time(rows_by_rank, *, window: int = 100, lookback_factor: int = 4,
                   include_series: bool = True, include_rank_heatmap: bool = False):
    """compute.py:129-315."""
    empty = lambda msg: {"metrics": [], "status_message": msg, "rank_heatmap": None}  # noqa: E731
    ranks = sorted(int(r) for r in rows_by_rank if len(rows_by_rank[r]))
    if not ranks:
        return empty("No ranks available")
    per = load_last_steps(rows_by_rank, window, lookback_factor)
    if not per:
        return empty("No StepTime data available")
    completed = min(max(m.keys()) for m in per.values() if m)
    steps = common_steps(per, completed, window)
    if not steps:
        return empty("No common step window yet")
    coverage = {"expected_steps": window, "steps_used": len(steps), "completed_step": int(completed),
                "world_size": len(ranks), "ranks_present": len(per),
                "incomplete": len(per) < len(ranks)}
    present = list(per.keys())
    sums: Dict[str, Dict[int, float]] = {k: {} for k in METRIC_KEYS}
    for rank, step_map in per.items():  # compute.py:502-531
        totals = {k: 0.0 for k in METRIC_KEYS}
        for st in steps:
            payload = step_map.get(int(st), {})
            for k in METRIC_KEYS:
                totals[k] += event_total_ms(payload, k)
        for k, t in totals.items():
            sums[k][int(rank)] = float(t)
    g = lambda k, r: sums.get(k, {}).get(r, 0.0)  # noqa: E731
    wait = {r: max(0.0, g("step_time", r) - g("h2d", r) - g("forward", r) - g("backward", r)
                   - g("optimizer_step", r)) for r in present}
    sums["wait_proxy"] = wait
    scores = {int(r): float(_sf(g("dataloader_fetch", r))
                            + max(_sf(g("step_time", r)),
                                  _sf(g("h2d", r)) + _sf(g("forward", r)) + _sf(g("backward", r))
                                  + _sf(g("optimizer_step", r)))) for r in present}
    worst_rank = max(scores, key=scores.get) if scores else None
    median_rank = None
    if scores:
        target = float(np.median(np.array(list(scores.values()), dtype=np.float64)))
        median_rank = min(scores, key=lambda r: (abs(scores[r] - target), scores[r], r))
    metrics: Dict[str, Any] = {}
    for k in METRIC_KEYS:
        m = make_metric(k, sums.get(k, {}), present, coverage, include_series, per, steps)
        if m is not None:
            if k == "step_time" and worst_rank is not None:
                m["summary"]["worst_rank"] = int(worst_rank)
            metrics[k] = m
    wm = make_metric("wait_proxy", wait, present, coverage, False, per, steps)
    if wm is not None:
        metrics["wait_proxy"] = wm
    heat = None
    if include_rank_heatmap and metrics:
        keys = [k for k in HEATMAP_KEYS if k in sums]
        rows = [{"rank": int(r), "sums_ms": {k: float(sums.get(k, {}).get(r, 0.0)) for k in keys}}
                for r in present]
        rows.sort(key=lambda row: (scores.get(row["rank"], 0.0), row["sums_ms"].get("step_time", 0.0),
                                   row["sums_ms"].get("dataloader_fetch", 0.0)), reverse=True)
        heat = {"window_size": window, "steps_used": coverage["steps_used"], "metric_keys": keys,
                "rows": rows, "sort_by": ["overall_score", "step_time", "dataloader_fetch"]}
    status = "OK"
    if worst_rank is not None:
        status += f" | overall_worst_rank=r{worst_rank}"
    if median_rank is not None:
        status += f" | overall_median_rank=r{median_rank}"
    return {"metrics": list(metrics.values()), "status_message": status, "rank_heatmap": heat}


# --------------------------------------------------------------------------- step memory
def live_step_memory(mem_rows_by_rank, *, window: int = 100, gpu_available: Optional[bool] = None,
                     metric_keys: Sequence[str] = ("peak_allocated", "peak_reserved")):
    """``build_step_memory_combined_result`` (renderers/step_memory/common.py:215-356).

    ``mem_rows_by_rank[rank]`` = ``[(step, peak_alloc | None, peak_resv | None), ...]`` in
    insertion order (id order).  ``device`` (a majority vote broken by Python set order,
    ``common.py:400-408``) is not reproduced.
    """
    ws = max(1, int(window))
    latest = {int(r): max(int(s) for s, _, _ in rows) for r, rows in mem_rows_by_rank.items() if len(rows)}
    if not latest:  # common.py:235-239
        return {"metrics": [], "status_message": "Waiting for first fully completed step across all ranks…"}
    world_size = len(latest)
    completed = min(latest.values())
    scan_span = max(ws * 20, ws + 1)  # common.py:245
    start = max(0, completed - scan_span + 1)
    rows_in_window = sum(1 for rows in mem_rows_by_rank.values() for s, _, _ in rows
                         if start <= int(s) <= completed)
    out = []
    for mi, key in enumerate(metric_keys):
        col = {"peak_allocated": 1, "peak_reserved": 2}[key]
        rank_maps: Dict[int, Dict[int, float]] = {}
        for r in sorted(mem_rows_by_rank):  # common.py:143-178: step DESC, id DESC, first wins
            rows = mem_rows_by_rank[r]
            cand = [(int(row[0]), i, row[col]) for i, row in enumerate(rows)
                    if row[col] is not None and start <= int(row[0]) <= completed]
            cand.sort(key=lambda t: (t[0], t[1]), reverse=True)
            m: Dict[int, float] = {}
            for s, _, v in cand:
                if len(m) >= scan_span:
                    break
                if s in m:
                    continue
                m[s] = float(v)
            if m:
                rank_maps[int(r)] = m
        if not rank_maps:
            continue
        maps = list(rank_maps.values())
        steps_rev: List[int] = []  # common.py:359-397
        step, scanned = int(completed), 0
        while step >= 0 and len(steps_rev) < ws:
            scanned += 1
            if scanned > scan_span:
                break
            if all(step in m for m in maps):
                steps_rev.append(step)
            step -= 1
        if not steps_rev:
            continue
        steps = steps_rev[::-1]
        ranks_list = sorted(rank_maps)
        values = np.array([[rank_maps[r][s] for s in steps] for r in ranks_list], dtype=np.float64)
        median_arr = np.median(values, axis=0)
        worst_arr = np.max(values, axis=0)
        peaks = np.max(values, axis=1)
        median_peak = float(np.median(peaks))
        worst_peak = float(np.max(peaks))
        worst_rank = int(ranks_list[int(np.argmax(peaks))])
        skew_ratio = (worst_peak / median_peak) if median_peak > 0.0 else 0.0
        skew_pct = ((worst_peak - median_peak) / median_peak) if median_peak > 0.0 else 0.0
        out.append({
            "metric": str(key),
            "series": {"steps": [int(s) for s in steps], "median": median_arr.astype(float).tolist(),
                       "worst": worst_arr.astype(float).tolist()},
            "summary": {"window_size": ws, "steps_used": len(steps), "median_peak": median_peak,
                        "worst_peak": worst_peak, "worst_rank": worst_rank,
                        "skew_ratio": float(skew_ratio), "skew_pct": float(skew_pct)},
            "coverage": {"expected_steps": ws, "steps_used": len(steps), "completed_step": int(completed),
                         "world_size": int(world_size), "ranks_present": len(ranks_list),
                         "incomplete": len(ranks_list) < world_size},
        })
    if out:
        status = "OK"
    elif gpu_available is False and rows_in_window > 0:
        status = "No GPU detected. Step memory uses torch-based GPU memory telemetry."
    else:
        status = "No complete memory metrics available"
    return {"metrics": out, "status_message": status}
