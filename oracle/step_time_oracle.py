"""Oracle: Step-Time reduce + diagnosis.  TEST INFRASTRUCTURE ONLY.

CPU restatement (plain Python + numpy, same operation order as the reference)
of the Step-Time half of the final-summary path:

  rows per rank -> buckets -> per-step metrics -> per-rank means
    (src/traceml/reporting/sections/step_time/model.py:50-74,162-281)
  common-suffix alignment + re-averaging
    (src/traceml/utils/step_windows.py:14-33,
     src/traceml/reporting/sections/step_time/alignment.py:44-155)
  cross-rank median / worst series and rank aggregates
    (src/traceml/diagnostics/step_time/adapters.py:92-355)
  context, rules, primary selection, trend note
    (src/traceml/diagnostics/step_time/context.py:84-532, rules.py:87-296,
     api.py:118-649, trend.py:36-147, policy.py:55-73)
  public rollups used by the kept payload builder
    (src/traceml/reporting/sections/step_time/model.py:77-105,382-498)

Inputs are the rows the reference's loader yields
(``loader.py:44-72``): for each global rank a list of
``{"step": int, "events": {name: {device: {"duration_ms": float, ...}}}}``
ordered ``step DESC, id DESC`` and already limited to ``max_rows``.
Outputs are plain dicts whose keys match the reference dataclass fields, so
they compare 1:1 with ``dataclasses.asdict`` of the reference objects.
"""

from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .trend_oracle import format_trend_pct, trend_pct

# policy.py:55-73 (SUMMARY_STEP_TIME_POLICY)
SUMMARY_THRESHOLDS = {
    "input_straggler_score_warn": 0.10,
    "input_straggler_score_crit": 0.18,
    "compute_straggler_score_warn": 0.10,
    "compute_straggler_score_crit": 0.18,
    "input_share_warn": 0.30,
    "input_share_crit": 0.40,
    "wait_share_warn": 0.18,
    "wait_share_crit": 0.28,
    "input_bound_max_skew": 0.05,
    "compute_bound_max_skew": 0.05,
    "compute_bound_share_warn": 0.88,
    "compute_bound_share_crit": 0.94,
    "min_steps_for_confident_diag": 20,
}
SUMMARY_MIN_STEPS_FOR_DIAG = 50

METRIC_KEYS = (
    "dataloader_fetch",
    "forward",
    "backward",
    "optimizer_step",
    "step_time",
    "wait_proxy",
)

STATUS_BY_KIND = {  # api.py:53-63
    "NO_DATA": "NO DATA",
    "WARMUP": "WARMUP",
    "BALANCED": "BALANCED",
    "STRAGGLER": "STRAGGLER",
    "INPUT_STRAGGLER": "INPUT STRAGGLER",
    "COMPUTE_STRAGGLER": "COMPUTE STRAGGLER",
    "INPUT_BOUND": "INPUT-BOUND",
    "COMPUTE_BOUND": "COMPUTE-BOUND",
    "WAIT_HEAVY": "WAIT-HEAVY",
}
PRIMARY_PRIORITY = {  # api.py:65-72
    "STRAGGLER": 50,
    "INPUT_STRAGGLER": 40,
    "COMPUTE_STRAGGLER": 39,
    "INPUT_BOUND": 30,
    "WAIT_HEAVY": 20,
    "COMPUTE_BOUND": 10,
}
SEVERITY_RANK = {"crit": 2, "warn": 1, "info": 0}  # diagnostics/common.py:98-102


# --------------------------------------------------------------------------
# scalars
# --------------------------------------------------------------------------
def ffloat(x: Any) -> float:
    """model.py:31-34 (+ summary_formatting.safe_float): non-finite -> 0.0."""
    try:
        v = float(x)
    except Exception:
        return 0.0
    return v if math.isfinite(v) else 0.0


def nnf(x: Any) -> float:
    """context.py:84-95: finite, clamped at >= 0."""
    try:
        v = float(x)
    except Exception:
        return 0.0
    if not math.isfinite(v):
        return 0.0
    return max(0.0, v)


def share(value: float, total: float) -> float:
    """context.py:98-105."""
    t = nnf(total)
    if t <= 0.0:
        return 0.0
    return max(0.0, nnf(value) / t)


def pct_str(v: float) -> str:
    return f"{nnf(v) * 100.0:.1f}%"


def rank_str(r: Optional[int]) -> str:
    return f"r{r}" if r is not None else "—"


# --------------------------------------------------------------------------
# a11: rows -> buckets -> per-rank analysis
# --------------------------------------------------------------------------
def bucket_of(name: str) -> Optional[str]:
    """model.py:50-74 -- substring bucketing (``h2d_time`` matches nothing)."""
    n = str(name).lower()
    for needle, bucket in (
        ("step_time", "step_time"),
        ("dataloader_next", "dataloader"),
        ("forward_time", "forward"),
        ("backward_time", "backward"),
        ("optimizer_step", "optimizer"),
    ):
        if needle in n:
            return bucket
    if "data" in n or "dataloader" in n or "input" in n or "batch" in n:
        return "dataloader"
    if "forward" in n or n == "fwd":
        return "forward"
    if "backward" in n or "bwd" in n:
        return "backward"
    if "optim" in n or "optimizer" in n or n in {"step", "update"}:
        return "optimizer"
    return None


def row_buckets(events: Dict[str, Any]) -> Optional[Dict[str, float]]:
    """model.py:37-47,162-200 -- sum duration_ms over devices per bucket."""
    m = {"dataloader": 0.0, "forward": 0.0, "backward": 0.0,
         "optimizer": 0.0, "step_time": 0.0}
    for name, by_dev in events.items():
        b = bucket_of(str(name))
        if b is None:
            continue
        tot = 0.0
        if isinstance(by_dev, dict):
            for stats in by_dev.values():
                if isinstance(stats, dict):
                    tot += ffloat(stats.get("duration_ms"))
        m[b] += tot
    if all(m[k] <= 0.0 for k in m):
        return None
    return m


def _summary(n, s_dl, s_f, s_b, s_o, s_cpu, s_tr, s_tot) -> Dict[str, Any]:
    return {
        "steps_analyzed": n,
        "avg_dataloader_ms": s_dl / n,
        "avg_forward_ms": s_f / n,
        "avg_backward_ms": s_b / n,
        "avg_optimizer_ms": s_o / n,
        "avg_step_cpu_ms": s_cpu / n,
        "avg_traced_step_ms": s_tr / n,
        "avg_gpu_compute_ms": (s_f + s_b + s_o) / n,
        "avg_total_step_ms": s_tot / n,
    }


def rank_analysis(rows: Sequence[Dict[str, Any]]):
    """model.py:203-281.  Returns (summary, per_step_metrics) or None.

    Sums run in row order; a later row with the same step id overwrites the
    per-step entry but both rows are counted in the means (as the reference).
    """
    if not rows:
        return None
    s_dl = s_f = s_b = s_o = s_cpu = s_tr = s_tot = 0.0
    n = 0
    per_step: Dict[int, Dict[str, float]] = {}
    for row in rows:
        step_id = row.get("step")
        m = row_buckets(row["events"])
        if m is None or step_id is None:
            continue
        dl, f, b, o = (ffloat(m["dataloader"]), ffloat(m["forward"]),
                       ffloat(m["backward"]), ffloat(m["optimizer"]))
        cpu = ffloat(m["step_time"])
        compute = f + b + o
        traced = max(cpu, compute)
        wait = max(0.0, traced - compute)
        per_step[int(step_id)] = {
            "dataloader_fetch": dl, "forward": f, "backward": b,
            "optimizer_step": o, "step_time": traced, "wait_proxy": wait,
        }
        s_dl += dl; s_f += f; s_b += b; s_o += o
        s_cpu += cpu; s_tr += traced; s_tot += dl + traced
        n += 1
    if n == 0:
        return None
    return _summary(n, s_dl, s_f, s_b, s_o, s_cpu, s_tr, s_tot), per_step


# --------------------------------------------------------------------------
# a12: alignment
# --------------------------------------------------------------------------
def common_suffix_steps(per_rank_steps: Dict[int, Dict[int, Any]],
                        max_rows: int) -> List[int]:
    """utils/step_windows.py:14-33."""
    if not per_rank_steps:
        return []
    sets = []
    for step_map in per_rank_steps.values():
        if not step_map:
            return []
        sets.append(set(int(s) for s in step_map.keys()))
    common = set.intersection(*sets) if sets else set()
    if not common:
        return []
    return sorted(common)[-max(1, int(max_rows)):]


def summary_from_step_metrics(step_metrics: Dict[int, Dict[str, float]]):
    """alignment.py:44-91 -- note ``step_time`` here is already the traced
    step, so avg_step_cpu_ms of an aligned summary is the traced mean."""
    if not step_metrics:
        return None
    s_dl = s_f = s_b = s_o = s_cpu = s_tr = s_tot = 0.0
    n = 0
    for m in step_metrics.values():
        dl = ffloat(m.get("dataloader_fetch"))
        f = ffloat(m.get("forward"))
        b = ffloat(m.get("backward"))
        o = ffloat(m.get("optimizer_step"))
        st = ffloat(m.get("step_time"))
        compute = f + b + o
        s_dl += dl; s_f += f; s_b += b; s_o += o
        s_cpu += max(0.0, st)
        traced = max(st, compute)
        s_tr += traced
        s_tot += dl + traced
        n += 1
    if n == 0:
        return None
    return _summary(n, s_dl, s_f, s_b, s_o, s_cpu, s_tr, s_tot)


def aligned_step_summary(per_rank_step_metrics, max_rows: int):
    """alignment.py:94-155 -> (aligned_summary, aligned_metrics, window)."""
    observed = len(per_rank_step_metrics)
    window_size = max(1, int(max_rows))
    common = common_suffix_steps(per_rank_step_metrics, max_rows)
    if not common:
        return {}, {}, {
            "alignment": "common_steps", "steps_analyzed": 0,
            "start_step": None, "end_step": None,
            "window_size": window_size, "global_ranks_used": 0,
            "global_ranks_observed": observed,
        }
    cset = set(common)
    a_metrics: Dict[int, Dict[int, Dict[str, float]]] = {}
    a_summary: Dict[int, Dict[str, Any]] = {}
    for rank, step_map in per_rank_step_metrics.items():
        rm = {int(s): m for s, m in step_map.items() if int(s) in cset}
        summ = summary_from_step_metrics(rm)
        if summ is None:
            continue
        a_metrics[int(rank)] = dict(sorted(rm.items()))
        a_summary[int(rank)] = summ
    return a_summary, a_metrics, {
        "alignment": "common_steps", "steps_analyzed": len(common),
        "start_step": int(common[0]), "end_step": int(common[-1]),
        "window_size": window_size, "global_ranks_used": len(a_summary),
        "global_ranks_observed": observed,
    }


def load_section(rows_by_rank: Dict[int, Sequence[Dict[str, Any]]],
                 max_rows: int, latest_step_observed: Optional[int]):
    """loader.py:109-184 minus SQL: the StepTimeSectionData numbers."""
    row_limit = max(1, int(max_rows))
    per_rank_summary, per_rank_steps = {}, {}
    for rank in sorted(rows_by_rank):
        res = rank_analysis(list(rows_by_rank[rank])[:row_limit])
        if res is not None:
            per_rank_summary[rank], per_rank_steps[rank] = res
    a_sum, a_met, window = aligned_step_summary(per_rank_steps, row_limit)
    return {
        "training_steps": (latest_step_observed + 1
                           if latest_step_observed is not None else 0),
        "latest_step_observed": latest_step_observed,
        "aligned_summary": a_sum,
        "aligned_step_metrics": a_met,
        "aligned_window": window,
        "per_global_rank_summary": per_rank_summary,
        "per_global_rank_step_metrics": per_rank_steps,
        "max_rows": row_limit,
    }


# --------------------------------------------------------------------------
# a13: cross-rank series + rank aggregates
# --------------------------------------------------------------------------
def metric_series(metric_key: str, steps: List[int], per_rank_step_metrics):
    """adapters.py:92-139 -- per step np.median / np.max over sorted ranks."""
    if not steps or not per_rank_step_metrics:
        return None
    ranks = sorted(per_rank_step_metrics.keys())
    med, worst = [], []
    for st in steps:
        vals = [ffloat(per_rank_step_metrics.get(r, {}).get(st, {})
                       .get(metric_key, 0.0)) for r in ranks]
        arr = np.asarray(vals, dtype=np.float64)
        if arr.size == 0:
            med.append(0.0); worst.append(0.0)
        else:
            med.append(float(np.median(arr)))
            worst.append(float(np.max(arr)))
    return {"steps": list(steps), "median": med, "worst": worst}


def metric_from_rank_values(metric_key, rank_values, coverage, series=None,
                            worst_rank_override=None):
    """adapters.py:142-197."""
    if not rank_values:
        return None
    ranks = sorted(int(r) for r in rank_values.keys())
    arr = np.asarray([ffloat(rank_values[r]) for r in ranks], dtype=np.float64)
    if arr.size == 0:
        return None
    median_total = float(np.median(arr))
    widx = int(np.argmax(arr))
    worst_total = float(arr[widx])
    worst_rank = int(ranks[widx])
    if coverage["ranks_present"] <= 1:
        median_total = worst_total
        skew_ratio = skew_pct = 0.0
    elif median_total > 0.0:
        skew_ratio = worst_total / median_total
        skew_pct = (worst_total - median_total) / median_total
    else:
        skew_ratio = skew_pct = 0.0
    if worst_rank_override is not None:
        worst_rank = int(worst_rank_override)
    return {
        "metric": str(metric_key),
        "series": series,
        "summary": {
            "window_size": int(coverage["expected_steps"]),
            "steps_used": int(coverage["steps_used"]),
            "median_total": median_total,
            "worst_total": worst_total,
            "worst_rank": worst_rank,
            "skew_ratio": float(skew_ratio),
            "skew_pct": float(skew_pct),
        },
        "coverage": coverage,
    }


def rank_signals_from_summary(aligned_summary):
    """model.py:142-159 (to_rank_signals)."""
    return {
        int(r): {
            "steps_analyzed": int(s["steps_analyzed"]),
            "dataloader_ms": ffloat(s["avg_dataloader_ms"]),
            "forward_ms": ffloat(s["avg_forward_ms"]),
            "backward_ms": ffloat(s["avg_backward_ms"]),
            "optimizer_ms": ffloat(s["avg_optimizer_ms"]),
            "step_cpu_ms": ffloat(s["avg_step_cpu_ms"]),
        }
        for r, s in aligned_summary.items()
    }


# --------------------------------------------------------------------------
# a14: context / rules / primary
# --------------------------------------------------------------------------
def _m_median(m):  # context.py:108-114
    return 0.0 if m is None else nnf(m["summary"]["median_total"])


def _m_worst(m):  # context.py:117-123
    return 0.0 if m is None else nnf(m["summary"]["worst_total"])


def _m_total(m, single):  # context.py:126-145
    if m is None:
        return 0.0
    return nnf(m["summary"]["worst_total"] if single
               else m["summary"]["median_total"])


def _m_skew(m, single):  # context.py:148-157
    if m is None or single:
        return 0.0
    return nnf(m["summary"]["skew_pct"])


def _m_wrank(m):  # context.py:160-171
    if m is None or m["summary"]["worst_rank"] is None:
        return None
    return int(m["summary"]["worst_rank"])


def _compute_candidates(fwd, bwd, opt, step_total, single):
    out = []
    for label, m in (("Forward", fwd), ("Backward", bwd), ("Optimizer", opt)):
        if m is None:
            continue
        total = _m_total(m, single)
        if total <= 0.0:
            continue
        out.append({"label": label, "share": share(total, step_total),
                    "skew": _m_skew(m, single), "worst_rank": _m_wrank(m)})
    return out


def build_context(metrics, th, per_rank_timing):
    """context.py:393-532."""
    by = {m["metric"]: m for m in metrics}
    step_m = by["step_time"]
    dl_m, wait_m = by.get("dataloader_fetch"), by.get("wait_proxy")
    fwd_m, bwd_m, opt_m = by.get("forward"), by.get("backward"), by.get("optimizer_step")
    cov = step_m["coverage"]
    single = (cov["world_size"] <= 1) or (cov["ranks_present"] <= 1)
    steps_used = int(step_m["summary"]["steps_used"])
    overall_worst = _m_wrank(step_m)
    step_total = _m_total(step_m, single)
    dl_total = _m_total(dl_m, single)
    wait_total = _m_total(wait_m, single)
    comp_total = (_m_total(fwd_m, single) + _m_total(bwd_m, single)
                  + _m_total(opt_m, single))
    cands = _compute_candidates(fwd_m, bwd_m, opt_m, step_total, single)
    # context.py:335,373 -- max() keeps the FIRST maximal candidate
    dominant = max(cands, key=lambda c: (c["skew"], c["share"])) if cands else None
    largest = max(cands, key=lambda c: c["share"]) if cands else None
    comp_skew = dominant["skew"] if dominant is not None else 0.0
    comp_rank = dominant["worst_rank"] if dominant is not None else overall_worst

    typical = _m_median(dl_m) + (_m_median(fwd_m) + _m_median(bwd_m) + _m_median(opt_m))
    if typical <= 0.0:
        in_score = comp_score = 0.0
    else:
        in_score = max(0.0, _m_worst(dl_m) - _m_median(dl_m)) / typical
        comp_score = max(
            0.0,
            (_m_worst(fwd_m) + _m_worst(bwd_m) + _m_worst(opt_m))
            - (_m_median(fwd_m) + _m_median(bwd_m) + _m_median(opt_m)),
        ) / typical

    local = {int(r): {str(k): nnf(v) for k, v in vals.items()}
             for r, vals in (per_rank_timing or {}).items()}
    if local:
        rank_values = {k: {r: nnf(v.get(k, 0.0)) for r, v in local.items()}
                       for k in METRIC_KEYS}
    else:  # context.py:376-391 (live fallback: worst-rank entry only)
        rank_values = {}
        for k, m in (("dataloader_fetch", dl_m), ("forward", fwd_m),
                     ("backward", bwd_m), ("optimizer_step", opt_m),
                     ("step_time", step_m), ("wait_proxy", wait_m)):
            r = _m_wrank(m)
            rank_values[k] = {} if (m is None or r is None) else {int(r): _m_worst(m)}
    return {
        "th": th, "single_rank": single, "steps_used": steps_used,
        "overall_worst_rank": overall_worst,
        "step_m": step_m, "dl_m": dl_m, "wait_m": wait_m,
        "fwd_m": fwd_m, "bwd_m": bwd_m, "opt_m": opt_m,
        "step_total": step_total, "dataloader_total": dl_total,
        "wait_total": wait_total, "compute_total": comp_total,
        "dataloader_share": share(dl_total, step_total),
        "wait_share": share(wait_total, step_total),
        "compute_share": share(comp_total, step_total),
        "dataloader_skew": _m_skew(dl_m, single),
        "compute_skew": comp_skew,
        "dataloader_worst_rank": _m_wrank(dl_m),
        "compute_worst_rank": comp_rank,
        "dominant_compute": dominant, "largest_compute": largest,
        "input_straggler_score": in_score,
        "compute_straggler_score": comp_score,
        "rank_values": rank_values, "per_rank_timing": local,
    }


def _sev(value, crit):
    return "crit" if nnf(value) >= crit else "warn"


def _issue(kind, status, severity, summary, action, metric=None, phase=None,
           score=None, share_pct=None, skew_pct=None, ranks=(), evidence=None):
    """rules.py:45-84 + diagnostics/common.py:44-67."""
    return {
        "kind": kind, "status": status, "severity": severity,
        "summary": summary, "action": action, "metric": metric, "phase": phase,
        "score": nnf(score) if score is not None else None,
        "share_pct": nnf(share_pct) if share_pct is not None else None,
        "skew_pct": nnf(skew_pct) if skew_pct is not None else None,
        "ranks": tuple(int(r) for r in ranks if r is not None),
        "evidence": dict(evidence or {}),
    }


def run_rules(c) -> List[Dict[str, Any]]:
    """rules.py:87-296 in registration order (:277-285)."""
    th, out = c["th"], []
    # InputStragglerRule :87-125
    if not c["single_rank"] and c["input_straggler_score"] >= th["input_straggler_score_warn"]:
        s, r = c["input_straggler_score"], c["dataloader_worst_rank"]
        out.append(_issue(
            "INPUT_STRAGGLER", "INPUT STRAGGLER",
            _sev(s, th["input_straggler_score_crit"]),
            f"{rank_str(r)} has excess dataloader burden "
            f"(~{pct_str(s)} of a typical local step).",
            f"Inspect input loading on {rank_str(r)}.",
            metric="dataloader_fetch", phase="dataloader", score=s,
            share_pct=c["dataloader_share"], skew_pct=c["dataloader_skew"],
            ranks=(r,)))
    # ComputeStragglerRule :128-174
    if not c["single_rank"] and c["compute_straggler_score"] >= th["compute_straggler_score_warn"]:
        s, r = c["compute_straggler_score"], c["compute_worst_rank"]
        label = c["dominant_compute"]["label"] if c["dominant_compute"] else "Compute"
        out.append(_issue(
            "COMPUTE_STRAGGLER", "COMPUTE STRAGGLER",
            _sev(s, th["compute_straggler_score_crit"]),
            f"{rank_str(r)} has excess compute burden "
            f"(~{pct_str(s)} of a typical local step).",
            f"Inspect {label.lower()} on {rank_str(r)}.",
            metric="compute", phase=label.lower(), score=s,
            share_pct=c["compute_share"], skew_pct=c["compute_skew"],
            ranks=(r,)))
    # InputBoundRule :177-212
    if c["dataloader_share"] >= th["input_share_warn"] and not (
            not c["single_rank"] and c["dataloader_skew"] > th["input_bound_max_skew"]):
        out.append(_issue(
            "INPUT_BOUND", "INPUT-BOUND",
            _sev(c["dataloader_share"], th["input_share_crit"]),
            f"Dataloader is {pct_str(c['dataloader_share'])} of the typical step.",
            "Increase workers, prefetch, or storage throughput.",
            metric="dataloader_fetch", phase="dataloader",
            share_pct=c["dataloader_share"], skew_pct=c["dataloader_skew"],
            ranks=(c["dataloader_worst_rank"],)))
    # WaitHeavyRule :215-246
    if c["wait_share"] >= th["wait_share_warn"]:
        out.append(_issue(
            "WAIT_HEAVY", "WAIT-HEAVY",
            _sev(c["wait_share"], th["wait_share_crit"]),
            f"WAIT* is {pct_str(c['wait_share'])} of the typical step.",
            "Inspect work outside traced phases, CPU stalls, logging, "
            "checkpointing, validation, or transfers.",
            metric="wait_proxy", phase="wait", share_pct=c["wait_share"],
            ranks=(c["overall_worst_rank"],)))
    # ComputeBoundRule :249-294
    if (c["compute_share"] >= th["compute_bound_share_warn"]
            and not c["dataloader_share"] >= th["input_share_warn"]
            and not c["wait_share"] >= th["wait_share_warn"]
            and not (not c["single_rank"]
                     and c["compute_skew"] > th["compute_bound_max_skew"])):
        label = c["largest_compute"]["label"] if c["largest_compute"] else "Compute"
        out.append(_issue(
            "COMPUTE_BOUND", "COMPUTE-BOUND",
            _sev(c["compute_share"], th["compute_bound_share_crit"]),
            f"Compute-bound; {label.lower()} is the largest phase.",
            "Optimize model compute or reduce step cost.",
            metric="compute", phase=label.lower(),
            share_pct=c["compute_share"], skew_pct=c["compute_skew"],
            ranks=(c["overall_worst_rank"],)))
    return out


def sort_issues(issues):
    """diagnostics/common.py:105-121 (stable, reverse=True)."""
    return sorted(
        issues,
        key=lambda i: (SEVERITY_RANK.get(i["severity"], 0),
                       float(i["score"] or 0.0), len(i["ranks"])),
        reverse=True)


def _diag(kind, severity, reason, action, steps_used, worst_rank=None, note=None):
    return {"kind": kind, "severity": severity, "status": STATUS_BY_KIND[kind],
            "reason": reason, "action": action, "steps_used": int(steps_used),
            "worst_rank": worst_rank, "note": note, "confidence": None}


def warmup_result(steps_used, required_steps, max_steps_used=None):
    """api.py:118-152."""
    low = max(0, int(steps_used))
    high = max(low, int(max_steps_used if max_steps_used is not None else low))
    required = max(1, int(required_steps))
    available = f"{low}" if low == high else f"{low}-{high}"
    suffix = "step" if high == 1 else "steps"
    primary = _diag(
        "WARMUP", "info",
        f"Only {available} {suffix} per rank available; summary "
        f"diagnosis requires {required}.",
        "Use a longer run for a stable timing diagnosis.", low)
    return {"primary": primary, "issues": (), "metric_attribution": {}}


def _series_trend(m, single):
    """trend.py:36-52."""
    if m is None or m.get("series") is None:
        return None
    s = m["series"]["worst"] if single else m["series"]["median"]
    if not s:
        return None
    return trend_pct(s)  # DEFAULT_TREND_CONFIG: 200 pts, 10% warm-up, 10k cap


def trend_note(kind, steps_used, single, step_m, wait_m, dl_m,
               wait_share, dl_share, wait_warn, input_warn):
    """trend.py:67-147 (min_steps 100, +-8 % state gates, 3 % dead-band)."""
    if steps_used < 100:
        return None

    def state(p):
        if p is None:
            return None
        if p >= 0.08:
            return "worsening"
        if p <= -0.08:
            return "improving"
        return None

    step_tr = _series_trend(step_m, single)
    wait_tr = _series_trend(wait_m, single)
    dl_tr = _series_trend(dl_m, single)
    ss, ws, ds = state(step_tr), state(wait_tr), state(dl_tr)
    fmt = lambda p: format_trend_pct(p, deadband_pct=0.03)  # noqa: E731
    if kind in {"INPUT_BOUND", "INPUT_STRAGGLER"} and ds:
        return f"Trend: dataloader is {ds} ({fmt(dl_tr)})."
    if kind in {"COMPUTE_BOUND", "COMPUTE_STRAGGLER", "STRAGGLER"} and ss:
        return f"Trend: step time is {ss} ({fmt(step_tr)})."
    if kind == "WAIT_HEAVY" and ws:
        return f"Trend: WAIT* is {ws} ({fmt(wait_tr)})."
    near_wait = wait_share >= wait_warn * 0.90
    near_input = dl_share >= input_warn * 0.90
    if kind == "BALANCED" and ss == "worsening" and (near_wait or near_input):
        return f"Trend: step time is rising ({fmt(step_tr)})."
    return None


def top_rank_entries(rank_values, max_items=3):
    """api.py:192-233 -- NB upper median ``values[n // 2]``."""
    if not rank_values:
        return []
    ordered = sorted(((int(r), nnf(v)) for r, v in rank_values.items()),
                     key=lambda it: (-it[1], it[0]))
    values = sorted(v for _, v in ordered)
    med = values[len(values) // 2]
    out = []
    for r, v in ordered[:max(1, int(max_items))]:
        ex = max(0.0, v - med)
        out.append({"rank": r, "value_ms": v, "excess_vs_median_ms": ex,
                    "pct_vs_median": (ex / med) if med > 0.0 else None})
    return out


def _attr(m, key, rank_values, step_total, single, phase):
    """api.py:236-260."""
    return {
        "metric": key, "phase": phase,
        "median_total_ms": _m_median(m), "worst_total_ms": _m_worst(m),
        "worst_rank": _m_wrank(m), "skew_pct": _m_skew(m, single),
        "share_pct": share(_m_total(m, single), step_total),
        "top_ranks": top_rank_entries(rank_values),
    }


def diagnosis_result(metrics, th, per_rank_timing=None):
    """api.py:313-649 (build_step_diagnosis_result)."""
    names = [m["metric"] for m in metrics]
    if len(names) != len(set(names)):
        return {"primary": _diag("NO_DATA", "info",
                                 "Duplicate metric keys in diagnosis input.",
                                 "Check upstream aggregation.", 0),
                "issues": (), "metric_attribution": {}}
    by = {m["metric"]: m for m in metrics}
    step_m = by.get("step_time")
    if step_m is None:
        return {"primary": _diag("NO_DATA", "info", "step_time metric is missing.",
                                 "Wait for the first complete window.", 0),
                "issues": (), "metric_attribution": {}}
    cov = step_m["coverage"]
    single = (cov["world_size"] <= 1) or (cov["ranks_present"] <= 1)
    steps_used = int(step_m["summary"]["steps_used"])
    overall_worst = _m_wrank(step_m)
    step_total = _m_total(step_m, single)
    if step_total <= 0.0:
        return {"primary": _diag("NO_DATA", "info", "No usable step-time data yet.",
                                 "Wait for the first complete window.",
                                 steps_used, overall_worst),
                "issues": (), "metric_attribution": {}}
    if steps_used < th["min_steps_for_confident_diag"]:
        res = warmup_result(steps_used, th["min_steps_for_confident_diag"])
        res["primary"]["worst_rank"] = overall_worst
        return res

    c = build_context(metrics, th, per_rank_timing)
    issue_list = run_rules(c)
    find = lambda k: next((i for i in issue_list if i["kind"] == k), None)  # noqa: E731
    in_issue, comp_issue = find("INPUT_STRAGGLER"), find("COMPUTE_STRAGGLER")
    in_s, comp_s = c["input_straggler_score"], c["compute_straggler_score"]
    both = in_issue is not None and comp_issue is not None
    if both:
        comb_rank = (c["dataloader_worst_rank"] if in_s >= comp_s
                     else c["compute_worst_rank"])
        comb_ranks = tuple(sorted({r for r in (c["dataloader_worst_rank"],
                                               c["compute_worst_rank"])
                                   if r is not None}))
        issue_list.append({
            "kind": "STRAGGLER", "status": "STRAGGLER",
            "severity": _sev(max(in_s, comp_s),
                             max(th["input_straggler_score_crit"],
                                 th["compute_straggler_score_crit"])),
            "summary": "Both input and compute are uneven across ranks.",
            "action": "Inspect the slowest rank and both dominant phases.",
            "metric": "step_time", "phase": "combined",
            "score": max(in_s, comp_s), "share_pct": None, "skew_pct": None,
            "ranks": comb_ranks or ((comb_rank,) if comb_rank is not None else ()),
            "evidence": {"input_score": in_s, "compute_score": comp_s},
        })
    issues = sort_issues(issue_list)
    ranked = sorted(issues, key=lambda i: (PRIMARY_PRIORITY.get(i["kind"], 0),
                                           float(i["score"] or 0.0)),
                    reverse=True)
    top = ranked[0] if ranked else None
    multi = lambda r: None if c["single_rank"] else r  # noqa: E731
    if both:
        dom = (c["dataloader_worst_rank"] if in_s >= comp_s
               else c["compute_worst_rank"])
        primary = _diag(
            "STRAGGLER",
            _sev(max(in_s, comp_s), max(th["input_straggler_score_crit"],
                                        th["compute_straggler_score_crit"])),
            "Both input and compute are uneven across ranks.",
            "Inspect the slowest rank and both dominant phases.",
            c["steps_used"], dom,
            f"Input score {pct_str(in_s)}, compute score {pct_str(comp_s)}.")
    elif in_issue is not None:
        primary = _diag("INPUT_STRAGGLER", in_issue["severity"], in_issue["summary"],
                        in_issue["action"], c["steps_used"], c["dataloader_worst_rank"],
                        f"Dataloader share is {pct_str(c['dataloader_share'])}.")
    elif comp_issue is not None:
        primary = _diag("COMPUTE_STRAGGLER", comp_issue["severity"],
                        comp_issue["summary"], comp_issue["action"], c["steps_used"],
                        c["compute_worst_rank"],
                        f"Compute share is {pct_str(c['compute_share'])}.")
    elif top is not None and top["kind"] == "INPUT_BOUND":
        primary = _diag("INPUT_BOUND", top["severity"], top["summary"], top["action"],
                        c["steps_used"], multi(c["dataloader_worst_rank"]))
    elif top is not None and top["kind"] == "WAIT_HEAVY":
        primary = _diag("WAIT_HEAVY", top["severity"], top["summary"], top["action"],
                        c["steps_used"], multi(c["overall_worst_rank"]),
                        "wait_ms = total_step_ms - dataloader_ms - compute_ms.")
    elif top is not None and top["kind"] == "COMPUTE_BOUND":
        primary = _diag("COMPUTE_BOUND", top["severity"], top["summary"], top["action"],
                        c["steps_used"], multi(c["overall_worst_rank"]))
    else:
        primary = _diag("BALANCED", "info",
                        "No dominant bottleneck is visible in this window.",
                        "Focus on throughput only if overall speed is still low.",
                        c["steps_used"], multi(c["overall_worst_rank"]))

    note = trend_note(primary["kind"], primary["steps_used"], c["single_rank"],
                      c["step_m"], c["wait_m"], c["dl_m"], c["wait_share"],
                      c["dataloader_share"], th["wait_share_warn"],
                      th["input_share_warn"])
    if note:  # api.py:155-160 (_merge_note)
        primary["note"] = note if not primary["note"] else f"{primary['note']} {note}"

    if not issues:
        issues = [_issue(primary["kind"], primary["status"], primary["severity"],
                         primary["reason"], primary["action"],
                         ranks=((primary["worst_rank"],)
                                if primary["worst_rank"] is not None else ()))]

    rv = c["rank_values"]
    fwd_rv, bwd_rv, opt_rv = rv.get("forward", {}), rv.get("backward", {}), rv.get("optimizer_step", {})
    comp_rv = {int(r): nnf(fwd_rv.get(r, 0.0)) + nnf(bwd_rv.get(r, 0.0))
               + nnf(opt_rv.get(r, 0.0))
               for r in sorted(set(fwd_rv) | set(bwd_rv) | set(opt_rv))}
    st, sg = c["step_total"], c["single_rank"]
    attribution = {
        "dataloader_fetch": _attr(c["dl_m"], "dataloader_fetch",
                                  rv.get("dataloader_fetch", {}), st, sg, "dataloader"),
        "forward": _attr(c["fwd_m"], "forward", fwd_rv, st, sg, "forward"),
        "backward": _attr(c["bwd_m"], "backward", bwd_rv, st, sg, "backward"),
        "optimizer_step": _attr(c["opt_m"], "optimizer_step", opt_rv, st, sg, "optimizer"),
        "wait_proxy": _attr(c["wait_m"], "wait_proxy", rv.get("wait_proxy", {}), st, sg, "wait"),
        "step_time": _attr(c["step_m"], "step_time", rv.get("step_time", {}), st, sg, "step"),
        "compute": {
            "metric": "compute",
            "phase": (c["dominant_compute"]["label"].lower()
                      if c["dominant_compute"] is not None else "compute"),
            "median_total_ms": _m_median(c["fwd_m"]) + _m_median(c["bwd_m"]) + _m_median(c["opt_m"]),
            "worst_total_ms": _m_worst(c["fwd_m"]) + _m_worst(c["bwd_m"]) + _m_worst(c["opt_m"]),
            "worst_rank": c["compute_worst_rank"],
            "skew_pct": c["compute_skew"], "share_pct": c["compute_share"],
            "top_ranks": top_rank_entries(comp_rv),
        },
    }
    return {"primary": primary, "issues": tuple(issues),
            "metric_attribution": attribution}


def diagnose_summary(rank_signals, *, max_rows, per_rank_step_metrics=None,
                     thresholds=None, min_steps_for_diag=SUMMARY_MIN_STEPS_FOR_DIAG,
                     return_metrics=False, precomputed=None):
    """adapters.py:232-355 (build_summary_step_diagnosis_result).

    ``precomputed`` (oracle/fast_oracle.py, large windows): ``{"steps_used", "completed_step",
    "series": {metric_key: {"steps", "median", "worst"}}}`` -- the per-step series already
    reduced with numpy instead of the per-step Python loop of ``metric_series``; everything
    downstream is this function unchanged."""
    th = dict(thresholds or SUMMARY_THRESHOLDS)
    if not rank_signals:
        return None
    ranks = sorted(rank_signals.keys())
    min_steps = min(s["steps_analyzed"] for s in rank_signals.values())
    max_steps = max(s["steps_analyzed"] for s in rank_signals.values())
    if min_steps < int(min_steps_for_diag):
        return warmup_result(int(min_steps), int(min_steps_for_diag), int(max_steps))
    common, series = [], {}
    if precomputed is not None:
        series = dict(precomputed["series"])
    elif per_rank_step_metrics:
        common = common_suffix_steps(per_rank_step_metrics, max_rows=max_rows)
        for mk in METRIC_KEYS:
            series[mk] = metric_series(mk, common, per_rank_step_metrics)
    n_common = int(precomputed["steps_used"]) if precomputed is not None else len(common)
    last_common = (int(precomputed["completed_step"]) if precomputed is not None
                   else (int(common[-1]) if common else 0))
    coverage = {
        "expected_steps": int(max_rows),
        "steps_used": n_common if n_common else int(min_steps),
        "completed_step": last_common if n_common else 0,
        "world_size": len(ranks), "ranks_present": len(ranks),
        "incomplete": False,
    }
    dl = {r: ffloat(s["dataloader_ms"]) for r, s in rank_signals.items()}
    fwd = {r: ffloat(s["forward_ms"]) for r, s in rank_signals.items()}
    bwd = {r: ffloat(s["backward_ms"]) for r, s in rank_signals.items()}
    opt = {r: ffloat(s["optimizer_ms"]) for r, s in rank_signals.items()}
    raw = {r: ffloat(s["step_cpu_ms"]) for r, s in rank_signals.items()}
    compute = {r: fwd[r] + bwd[r] + opt[r] for r in ranks}
    eff = {r: max(raw[r], compute[r]) for r in ranks}
    wait = {r: max(0.0, eff[r] - compute[r]) for r in ranks}
    overall = {r: dl[r] + eff[r] for r in ranks}
    overall_worst = int(max(ranks, key=lambda r: (overall[r], -r)))
    values = {"dataloader_fetch": dl, "forward": fwd, "backward": bwd,
              "optimizer_step": opt, "step_time": eff, "wait_proxy": wait}
    metrics = []
    for key in METRIC_KEYS:
        m = metric_from_rank_values(
            key, values[key], coverage, series.get(key),
            overall_worst if key == "step_time" else None)
        if m is not None:
            metrics.append(m)
    if not metrics:
        return None
    # adapters.py:200-229 (_build_summary_per_rank_timing)
    prt = {}
    for r, s in rank_signals.items():
        f, b, o = ffloat(s["forward_ms"]), ffloat(s["backward_ms"]), ffloat(s["optimizer_ms"])
        c = f + b + o
        e = max(ffloat(s["step_cpu_ms"]), c)
        prt[int(r)] = {"dataloader_fetch": ffloat(s["dataloader_ms"]), "forward": f,
                       "backward": b, "optimizer_step": o, "step_time": e,
                       "wait_proxy": max(0.0, e - c)}
    res = diagnosis_result(metrics, th, prt)
    if return_metrics:
        res = dict(res)
        res["_metrics"] = metrics
    return res


# --------------------------------------------------------------------------
# a15: public rollups consumed by the kept payload builder
# --------------------------------------------------------------------------
def closest_rank_to_median(rank_to_value: Dict[int, float]) -> Optional[int]:
    """model.py:77-105 (tie: |d|, value, rank)."""
    if not rank_to_value:
        return None
    vals = np.asarray([ffloat(v) for v in rank_to_value.values()], dtype=np.float64)
    if vals.size == 0:
        return None
    med = float(np.median(vals))
    return min(rank_to_value.keys(),
               key=lambda r: (abs(ffloat(rank_to_value[r]) - med),
                              ffloat(rank_to_value[r]), r))


def wait_avg_ms(s) -> float:
    """model.py:284-307."""
    return max(0.0, ffloat(s["avg_traced_step_ms"])
               - (ffloat(s["avg_forward_ms"]) + ffloat(s["avg_backward_ms"])
                  + ffloat(s["avg_optimizer_ms"])))


def rank_metric_values(summary_by_rank):
    """model.py:310-344."""
    g = lambda key: {int(r): ffloat(s[key]) for r, s in summary_by_rank.items()}  # noqa: E731
    return {
        "total_step_ms": g("avg_total_step_ms"),
        "dataloader_ms": g("avg_dataloader_ms"),
        "compute_ms": g("avg_gpu_compute_ms"),
        "wait_ms": {int(r): wait_avg_ms(s) for r, s in summary_by_rank.items()},
        "forward_ms": g("avg_forward_ms"),
        "backward_ms": g("avg_backward_ms"),
        "optimizer_ms": g("avg_optimizer_ms"),
    }


def global_points(summary_by_rank):
    """model.py:361-403,406-442: average / median{value,idx} / worst{value,idx}."""
    vbm = rank_metric_values(summary_by_rank)
    avg, med, worst = {}, {}, {}
    for metric, values in vbm.items():
        vs = list(values.values())
        avg[metric] = sum(vs) / len(vs) if vs else None
        if not values:
            med[metric] = worst[metric] = {"value": None, "idx": None}
            continue
        mr = closest_rank_to_median(values)
        wr = max(values, key=lambda r: (values[r], -int(r)))
        med[metric] = {"value": values.get(mr), "idx": str(mr)}
        worst[metric] = {"value": values.get(wr), "idx": str(wr)}
    return {"average": avg, "median": med, "worst": worst}


def overview(summary_by_rank):
    """model.py:445-498."""
    if not summary_by_rank:
        return {"rank_comparison": "no_data", "median_global_rank": None,
                "worst_global_rank": None, "median_avg_step_ms": None,
                "worst_avg_step_ms": None, "step_time_skew_percent": None}
    tot = {r: s["avg_total_step_ms"] for r, s in summary_by_rank.items()}
    wr = max(tot, key=tot.get)
    mr = closest_rank_to_median(tot)
    w, m = tot.get(wr), (tot.get(mr) if mr is not None else None)
    skew = None
    if w is not None and m is not None and m > 0.0 and wr != mr:
        skew = 100.0 * (w - m) / m
    return {"rank_comparison": "single_rank" if len(summary_by_rank) <= 1 else "distributed",
            "median_global_rank": mr, "worst_global_rank": wr,
            "median_avg_step_ms": m, "worst_avg_step_ms": w,
            "step_time_skew_percent": skew}


def step_time_section(rows_by_rank, *, max_rows=10_000, latest_step_observed=None):
    """Whole Step-Time path: load -> diagnose (sections/step_time/__init__.py:42-94)."""
    if latest_step_observed is None:
        steps = [r["step"] for rows in rows_by_rank.values() for r in rows
                 if r.get("step") is not None]
        latest_step_observed = max(steps) if steps else None
    data = load_section(rows_by_rank, max_rows, latest_step_observed)
    diag = diagnose_summary(
        rank_signals_from_summary(data["aligned_summary"]),
        max_rows=data["max_rows"],
        per_rank_step_metrics=data["aligned_step_metrics"])
    return {"data": data, "diagnosis": diag,
            "global": global_points(data["aligned_summary"]),
            "overview": overview(data["aligned_summary"])}
