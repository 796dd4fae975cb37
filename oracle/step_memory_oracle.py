"""Oracle: Step-Memory reduce + diagnosis.  TEST INFRASTRUCTURE ONLY.

Restates, on the CPU and in the reference's operation order:

  candidate rows / dedupe / common-suffix window
    (src/traceml/reporting/sections/step_memory/loader.py:112-254)
  per-step cross-rank median (mean of the two middles) / max, rank peaks,
  skew, window means
    (src/traceml/reporting/sections/step_memory/model.py:130-246)
  creep evidence
    (src/traceml/diagnostics/step_memory/trend.py:203-277 on
     src/traceml/analytics/trends/core.py:51-115 with min_points 50,
     warm-up 0 -- policy.py:31-36)
  signals, rules, ordering, primary
    (src/traceml/diagnostics/step_memory/adapters.py:72-172,
     rules.py:96-283, api.py:284-508, policy.py:12-36)
  public rollup points
    (src/traceml/reporting/sections/step_memory/model.py:322-412)

Input: ``rows_by_rank[rank] = [(step, peak_alloc_bytes, peak_resv_bytes), ...]``
in insertion (SQLite id) order; rows whose peaks are None are skipped as in
``loader.py:136-141``.
"""

from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Sequence, Tuple

from .step_time_oracle import SEVERITY_RANK, common_suffix_steps
from .trend_oracle import trend_evidence

TH = {  # policy.py:12-36
    "min_steps_for_diag": 50,
    "pressure_warn_fraction": 0.92,
    "pressure_crit_fraction": 0.97,
    "imbalance_skew_warn": 0.12,
    "imbalance_skew_crit": 0.20,
    "creep_score_delta_scale_bytes": 100.0 * 1024.0 * 1024.0,
    "creep_confirmed_delta_bytes": 1024.0 * 1024.0 * 1024.0,
}
STATUS = {  # api.py:36-43
    "NO_DATA": "NO DATA", "BALANCED": "BALANCED",
    "HIGH_PRESSURE": "HIGH PRESSURE", "IMBALANCE": "IMBALANCE",
    "CREEP_EARLY": "MEMORY RISING", "CREEP_CONFIRMED": "MEMORY CREEP",
}
ISSUE_PRIORITY = {"HIGH_PRESSURE": 0, "IMBALANCE": 1,
                  "CREEP_CONFIRMED": 2, "CREEP_EARLY": 3}  # rules.py:227-232


def median2(values: Sequence[float]) -> float:
    """model.py:130-138 -- sorted; even count -> mean of the two middles."""
    o = sorted(float(v) for v in values)
    if not o:
        return 0.0
    mid = len(o) // 2
    return o[mid] if len(o) % 2 else (o[mid - 1] + o[mid]) / 2.0


def candidate_rows(rows_by_rank, window_size: int):
    """loader.py:112-205: latest row per (rank, step), then the most recent
    ``max(20 W, W + 1)`` steps per rank."""
    limit = max(int(window_size) * 20, int(window_size) + 1)
    out: Dict[int, Dict[int, Tuple[float, float]]] = {}
    for rank, rows in rows_by_rank.items():
        latest: Dict[int, Tuple[float, float]] = {}
        for step, alloc, resv in rows:  # id ASC: later rows overwrite
            if step is None or alloc is None or resv is None:
                continue
            latest[int(step)] = (float(alloc), float(resv))
        keep = sorted(latest)[-limit:]
        if keep:
            out[int(rank)] = {s: latest[s] for s in keep}
    return out


def aligned_window(rows_by_rank, window_size: int):
    """loader.py:208-254."""
    seen = sum(1 for rows in rows_by_rank.values()
               if any(r[0] is not None for r in rows))
    cands = candidate_rows(rows_by_rank, window_size)
    common = common_suffix_steps(cands, max_rows=window_size)
    if not common:
        return {"steps": (), "per_global_rank": {}, "window_size": int(window_size),
                "global_ranks_seen": seen}
    per_rank = {}
    for rank, step_rows in sorted(cands.items()):
        al = {s: step_rows[s] for s in common if s in step_rows}
        if len(al) != len(common):
            continue
        per_rank[int(rank)] = dict(sorted(al.items()))
    return {"steps": tuple(int(s) for s in common), "per_global_rank": per_rank,
            "window_size": int(window_size), "global_ranks_seen": seen}


def combined_metrics(win) -> List[Dict[str, Any]]:
    """model.py:141-221 (allocated first, reserved second)."""
    steps = list(win["steps"])
    ranks = sorted(win["per_global_rank"].keys())
    if not steps or not ranks:
        return []
    used = len(win["per_global_rank"])
    out = []
    for idx, name in enumerate(("peak_allocated", "peak_reserved")):
        by_rank = [[float(win["per_global_rank"][r][s][idx]) for s in steps]
                   for r in ranks]
        cols = list(zip(*by_rank))
        med_series = [float(median2(c)) for c in cols]
        worst_series = [float(max(float(v) for v in c)) for c in cols]
        peaks = [max(v) for v in by_rank]
        med_peak = float(median2(peaks))
        worst_peak = float(max(peaks))
        worst_rank = int(ranks[peaks.index(worst_peak)])
        skew_ratio = worst_peak / med_peak if med_peak > 0.0 else 0.0
        skew_pct = (worst_peak - med_peak) / med_peak if med_peak > 0.0 else 0.0
        out.append({
            "metric": name,
            "series": {"steps": [int(s) for s in steps],
                       "median": med_series, "worst": worst_series},
            "summary": {"window_size": int(win["window_size"]),
                        "steps_used": len(steps), "median_peak": med_peak,
                        "worst_peak": worst_peak, "worst_rank": worst_rank,
                        "skew_ratio": float(skew_ratio), "skew_pct": float(skew_pct)},
            "coverage": {"expected_steps": int(win["window_size"]),
                         "steps_used": len(steps),
                         "completed_step": steps[-1] if steps else None,
                         "world_size": int(win["global_ranks_seen"]),
                         "ranks_present": used,
                         "incomplete": used < win["global_ranks_seen"]},
        })
    return out


def rank_means(win) -> Dict[str, Dict[str, float]]:
    """model.py:224-246 -- JSON rows are window MEANS per rank."""
    out = {}
    for rank, rows in sorted(win["per_global_rank"].items()):
        vals = list(rows.values())
        if not vals:
            continue
        out[str(rank)] = {
            "peak_allocated_bytes": sum(v[0] for v in vals) / len(vals),
            "peak_reserved_bytes": sum(v[1] for v in vals) / len(vals),
        }
    return out


def _nn_list(values):
    out = []
    for v in values:
        try:
            x = float(v)
        except Exception:
            x = 0.0
        if not math.isfinite(x):
            x = 0.0
        out.append(max(0.0, x))
    return out


def window_creep(worst_series, median_series, steps_used: int, th=TH):
    """trend.py:203-277."""
    empty = {"eligible": False, "baseline_avg_bytes": None, "mid_avg_bytes": None,
             "recent_avg_bytes": None, "overall_abs_delta_bytes": None,
             "overall_worst_growth_pct": None, "overall_median_growth_pct": None,
             "trend_window_steps": None, "avg_growth_bytes_per_step": None,
             "early": False, "confirmed": False, "score": 0.0}
    if int(steps_used) < int(th["min_steps_for_diag"]):
        return empty
    worst, median = _nn_list(worst_series), _nn_list(median_series)
    kw = dict(min_points=50, warmup_frac=0.0, history_limit=10_000)
    w_ev, m_ev = trend_evidence(worst, **kw), trend_evidence(median, **kw)
    if w_ev is None or m_ev is None:
        return empty
    abs_delta = float(w_ev["delta_vs_baseline"])
    w_growth, m_growth = w_ev["delta_pct_vs_baseline"], m_ev["delta_pct_vs_baseline"]
    recent_gt_mid = w_ev["delta_vs_mid"] > 0.0 and m_ev["delta_vs_mid"] > 0.0
    mid_gt_base = (w_ev["mid_avg"] > w_ev["baseline_avg"]
                   and m_ev["mid_avg"] > m_ev["baseline_avg"])
    ok = recent_gt_mid and mid_gt_base
    early = bool(ok and abs_delta > 0.0)
    confirmed = bool(ok and abs_delta >= float(th["creep_confirmed_delta_bytes"]))
    score = (max(0.0, abs_delta) / max(1.0, float(th["creep_score_delta_scale_bytes"]))
             + max(0.0, float(w_growth or 0.0)) * 10.0
             + max(0.0, float(m_growth or 0.0)) * 6.0)
    tws = min(len(worst), 1000)
    growth = None
    if tws >= 2:
        tail = worst[-tws:]
        growth = float(tail[-1] - tail[0]) / float(tws - 1)
    return {"eligible": True, "baseline_avg_bytes": w_ev["baseline_avg"],
            "mid_avg_bytes": w_ev["mid_avg"], "recent_avg_bytes": w_ev["recent_avg"],
            "overall_abs_delta_bytes": abs_delta,
            "overall_worst_growth_pct": w_growth,
            "overall_median_growth_pct": m_growth,
            "trend_window_steps": tws, "avg_growth_bytes_per_step": growth,
            "early": early, "confirmed": confirmed, "score": score}


def _label(name: str) -> str:
    return name.replace("_", " ")


def _creep_note(trend) -> Optional[str]:
    """rules.py:37-57."""
    parts = []
    if trend["baseline_avg_bytes"] is not None and trend["recent_avg_bytes"] is not None:
        parts.append(f"baseline {trend['baseline_avg_bytes']:.0f} B -> "
                     f"recent {trend['recent_avg_bytes']:.0f} B")
    if trend["overall_abs_delta_bytes"] is not None:
        d = trend["overall_abs_delta_bytes"]
        parts.append(f"{'+' if d >= 0.0 else '-'}{abs(d):.0f} B")
    if trend["overall_worst_growth_pct"] is not None:
        parts.append(f"(~{trend['overall_worst_growth_pct'] * 100.0:.0f}%)")
    return ", ".join(parts) if parts else None


def metric_signals(metric, gpu_total_bytes, th=TH):
    """adapters.py:133-172."""
    s, cov = metric["summary"], metric["coverage"]
    nn = lambda v: max(0.0, float(v)) if v is not None else 0.0  # noqa: E731
    worst = nn(s["worst_peak"])
    total = float(gpu_total_bytes) if gpu_total_bytes is not None else 0.0
    pressure = None if total <= 0.0 else max(0.0, float(worst) / total)
    creep = window_creep(metric["series"]["worst"], metric["series"]["median"],
                         int(s["steps_used"] or 0), th)
    trend = {k: creep[k] for k in (
        "eligible", "baseline_avg_bytes", "mid_avg_bytes", "recent_avg_bytes",
        "overall_abs_delta_bytes", "overall_worst_growth_pct",
        "overall_median_growth_pct", "early", "confirmed", "score")}
    return {
        "metric": metric["metric"], "device": metric.get("device"),
        "steps_used": int(s["steps_used"] or 0),
        "window_size": int(s["window_size"] or 0),
        "completed_step": int(cov["completed_step"] or 0),
        "ranks_seen": int(cov["ranks_present"] or 0),
        "worst_rank": s["worst_rank"],
        "worst_peak_bytes": worst, "median_peak_bytes": nn(s["median_peak"]),
        "skew_ratio": nn(s["skew_ratio"]), "skew_pct": nn(s["skew_pct"]),
        "pressure_frac": pressure, "trend": trend,
    }


def _mem_issue(kind, status, severity, summary, action, sig, score, evidence):
    """rules.py:66-93."""
    return {"kind": kind, "status": status, "severity": severity,
            "summary": summary, "action": action, "metric": sig["metric"],
            "phase": "memory",
            "score": float(score) if score is not None else None,
            "share_pct": None, "skew_pct": sig["skew_pct"],
            "ranks": ((int(sig["worst_rank"]),) if sig["worst_rank"] is not None else ()),
            "evidence": dict(evidence or {})}


def sort_mem_issues(issues):
    """rules.py:235-254."""
    return sorted(issues, key=lambda i: (
        ISSUE_PRIORITY.get(i["kind"], 100), -SEVERITY_RANK.get(i["severity"], 0),
        -float(i["score"] or 0.0), str(i["metric"] or "")))


def run_mem_rules(sig, th=TH):
    """rules.py:96-224,257-272."""
    out = []
    p = sig["pressure_frac"]
    ready = sig["steps_used"] >= int(th["min_steps_for_diag"])
    if p is not None and ready and p >= th["pressure_warn_fraction"]:
        out.append(_mem_issue(
            "HIGH_PRESSURE", "HIGH PRESSURE",
            "crit" if p >= th["pressure_crit_fraction"] else "warn",
            f"{_label(sig['metric'])} is near device capacity (~{p * 100.0:.0f}%).",
            "Reduce memory load.", sig, p, {"pressure_frac": p}))
    k = sig["skew_pct"]
    if ready and not k < th["imbalance_skew_warn"]:
        out.append(_mem_issue(
            "IMBALANCE", "IMBALANCE",
            "crit" if k >= th["imbalance_skew_crit"] else "warn",
            f"{_label(sig['metric'])} shows +{k * 100.0:.1f}% cross-rank skew.",
            "Inspect per-rank workload.", sig, k, {"skew_pct": k}))
    t = sig["trend"]
    ev = {"overall_abs_delta_bytes": t["overall_abs_delta_bytes"],
          "overall_worst_growth_pct": t["overall_worst_growth_pct"],
          "overall_median_growth_pct": t["overall_median_growth_pct"],
          "note": _creep_note(t)}
    if t["confirmed"]:
        out.append(_mem_issue(
            "CREEP_CONFIRMED", "MEMORY CREEP", "warn",
            f"{_label(sig['metric'])} is rising across the window.",
            "Check retained tensors or caches.", sig, t["score"], ev))
    if t["early"] and not t["confirmed"]:
        out.append(_mem_issue(
            "CREEP_EARLY", "MEMORY RISING", "info",
            f"{_label(sig['metric'])} is rising from early to recent steps.",
            "Watch the next window.", sig, t["score"], ev))
    return sort_mem_issues(out)


def _mk(kind, severity, metric, steps_used, reason, action, worst_rank=None,
        note=None, confidence=None):
    return {"kind": kind, "severity": severity, "status": STATUS[kind],
            "reason": reason, "action": action, "metric": metric,
            "steps_used": int(steps_used), "worst_rank": worst_rank,
            "note": note, "confidence": confidence}


def diagnose_summary(metrics, gpu_total_bytes=None, th=TH):
    """api.py:284-508 (build_step_memory_summary_diagnosis_result)."""
    issues, attribution = [], {}
    for m in metrics:
        sig = metric_signals(m, gpu_total_bytes, th)
        issues.extend(run_mem_rules(sig, th))
        attribution[m["metric"]] = sig
    issues = sort_mem_issues(issues)
    if issues:  # api.py:396-433
        top = issues[0]
        metric = str(top["metric"] or "peak_reserved")
        sig = attribution.get(metric, {})
        ranks = tuple(top["ranks"] or ())
        conf = {"HIGH_PRESSURE": 0.9 if top["severity"] == "crit" else 0.8,
                "IMBALANCE": 0.85 if top["severity"] == "crit" else 0.75,
                "CREEP_CONFIRMED": 0.88, "CREEP_EARLY": 0.60}.get(top["kind"])
        raw_note = top["evidence"].get("note")
        primary = _mk(top["kind"], top["severity"], metric,
                      int(sig.get("steps_used") or 0), top["summary"], top["action"],
                      int(ranks[0]) if ranks else sig.get("worst_rank"),
                      str(raw_note) if raw_note else None, conf)
    elif not metrics:  # api.py:444-453
        primary = _mk("NO_DATA", "info", "peak_reserved", 0,
                      "No step-memory data yet.", "Wait for more completed steps.",
                      confidence=0.0)
    else:  # api.py:463-497
        sigs = list(attribution.values())
        ready = [s for s in sigs if int(s["steps_used"] or 0) >= int(th["min_steps_for_diag"])]
        if not ready:
            best = max(sigs, key=lambda s: int(s["steps_used"] or 0))
            primary = _mk("NO_DATA", "info", str(best["metric"] or "peak_reserved"),
                          int(best["steps_used"] or 0),
                          f"Need at least {int(th['min_steps_for_diag'])} completed steps.",
                          "Keep monitoring.", best["worst_rank"], confidence=0.0)
        else:
            by = {str(s["metric"]): s for s in ready}
            base = by.get("peak_reserved") or by.get("peak_allocated") or ready[0]
            primary = _mk("BALANCED", "info", str(base["metric"] or "peak_reserved"),
                          int(base["steps_used"] or 0),
                          "No clear pressure, imbalance, or creep signal.",
                          "Keep monitoring.", base["worst_rank"], confidence=0.75)
    return {"primary": primary, "issues": tuple(issues),
            "metric_attribution": attribution}


def rollup_points(per_rank_means):
    """model.py:322-412: average + median/worst {value, idx} over the rows."""
    names = ("peak_allocated_bytes", "peak_reserved_bytes")
    avg, med, worst = {}, {}, {}
    for name in sorted(names):
        values = {k: float(v[name]) for k, v in per_rank_means.items()
                  if v.get(name) is not None and math.isfinite(float(v[name]))}
        avg[name] = sum(values.values()) / len(values) if values else None
        if not values:
            med[name] = worst[name] = {"value": None, "idx": None}
            continue
        o = sorted(values.values())
        mid = len(o) // 2
        mv = o[mid] if len(o) % 2 else (o[mid - 1] + o[mid]) / 2.0
        rk = lambda k: int(k) if str(k).lstrip("-").isdigit() else 0  # noqa: E731
        mk = min(values, key=lambda k: (abs(values[k] - mv), values[k], rk(k)))
        wk = max(values, key=lambda k: (values[k], -rk(k)))
        med[name] = {"value": values.get(mk), "idx": mk}
        worst[name] = {"value": values.get(wk), "idx": wk}
    return {"average": avg, "median": med, "worst": worst}


def step_memory_section(rows_by_rank, *, window_size=10_000, gpu_total_bytes=None):
    """sections/step_memory/__init__.py:45-94 minus SQL and payload building."""
    win = aligned_window(rows_by_rank, max(1, int(window_size)))
    metrics = combined_metrics(win)
    means = rank_means(win)
    steps = [r[0] for rows in rows_by_rank.values() for r in rows if r[0] is not None]
    latest = max(steps) if steps else None
    return {
        "training_steps": latest + 1 if latest is not None else 0,
        "latest_step_observed": latest,
        "window": {"steps": win["steps"], "window_size": win["window_size"],
                   "global_ranks_seen": win["global_ranks_seen"],
                   "global_ranks_used": len(win["per_global_rank"])},
        "metrics": metrics, "per_global_rank": means,
        "diagnosis": diagnose_summary(metrics, gpu_total_bytes),
        "global": rollup_points(means),
    }
