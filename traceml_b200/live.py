"""Live tick: the B200 twin of the reference's ``StepCombinedComputer``.

The reference's live CLI / dashboard recompute, every render tick, the combined
step-time view from SQLite (``src/traceml/renderers/step_time/compute.py:38-315``):
load the last ``4 x window`` rows per rank, intersect the step ids, sum each phase
over the last ``window`` common steps per rank, and build median / worst / sum series.
Here every rank's rows never leave its HBM ring: the tick is four small kernels per
rank on a side stream (``csrc/tml_combined.cuh``), one MIN all-reduce of the presence
bytes, two small vector all-gathers and one row all-gather over NVLink.  Nothing
touches the training stream.

Same class name, constructor meaning, ``compute_cli`` / ``compute_dashboard`` entry
points, stale handling and result shape (``asdict(StepCombinedTimeResult)``) as the
reference; only the data source differs (engines + a comm instead of a db path).
"""

from __future__ import annotations

import time
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

from .reduce import LocalComm

# renderers/step_time/schema.py / compute.py:25-36 -- phase order of the 64-B row
DEFAULT_METRIC_KEYS = ("dataloader_fetch", "h2d", "forward", "backward", "optimizer_step", "step_time")
DEFAULT_HEATMAP_KEYS = ("dataloader_fetch", "h2d", "forward", "backward", "optimizer_step",
                        "wait_proxy", "step_time")
_COL = {k: i for i, k in enumerate(DEFAULT_METRIC_KEYS)}
_INFO_LEN = 5
_ALIGN_LEN = 8


def _empty(msg: str) -> Dict[str, Any]:
    return {"metrics": [], "status_message": msg, "rank_heatmap": None}


class StepCombinedComputer:
    """compute.py:38-128 (constructor, ``compute_cli``, ``compute_dashboard``, ``_compute``)."""

    def __init__(self, engines: Sequence[Any], comm: Any = None, *, window_size: int = 100,
                 metric_keys: Optional[Sequence[str]] = None,
                 heatmap_keys: Optional[Sequence[str]] = None,
                 stale_ttl_s: Optional[float] = 30.0, lookback_factor: int = 4,
                 device: Optional[torch.device] = None, use_side_stream: bool = True) -> None:
        self.engines = list(engines)
        self.comm = comm or LocalComm()
        self.L = len(self.engines)
        self.window_size = max(1, int(window_size))
        self.metric_keys = list(metric_keys) if metric_keys is not None else list(DEFAULT_METRIC_KEYS)
        self.heatmap_keys = list(heatmap_keys) if heatmap_keys is not None else list(DEFAULT_HEATMAP_KEYS)
        unknown = [k for k in self.metric_keys if k not in _COL]
        if unknown:
            raise ValueError(f"unknown step-time metric keys: {unknown}")
        self.lookback_factor = max(1, int(lookback_factor))
        self.device = device or torch.device("cuda", self.engines[0].device)
        self._stream = None
        if use_side_stream and self.device.type == "cuda":
            self._stream = torch.cuda.Stream(device=self.device)
        self._last_ok: Optional[Dict[str, Any]] = None
        self._last_ok_ts = 0.0
        self._stale_ttl_s = float(stale_ttl_s) if stale_ttl_s is not None else None
        self.last_timings_ms: Dict[str, float] = {}

    # ------------------------------------------------------------------ public API
    def compute_cli(self) -> Dict[str, Any]:
        return self._compute(include_series=True, include_rank_heatmap=False)

    def compute_dashboard(self) -> Dict[str, Any]:
        return self._compute(include_series=False, include_rank_heatmap=True)

    def _compute(self, *, include_series: bool, include_rank_heatmap: bool) -> Dict[str, Any]:
        """compute.py:103-123.  The tick is a collective: an exception on one rank is
        an exception on all (the native calls fail identically or not at all)."""
        try:
            if self._stream is not None:
                with torch.cuda.stream(self._stream):
                    result = self._compute_impl(include_series=include_series,
                                                include_rank_heatmap=include_rank_heatmap)
            else:
                result = self._compute_impl(include_series=include_series,
                                            include_rank_heatmap=include_rank_heatmap)
        except Exception as exc:  # noqa: BLE001 -- the live view must never take training down
            return self._stale_or_empty(f"STALE (exception: {type(exc).__name__})")
        if not result["metrics"]:
            return self._stale_or_empty("STALE (no metrics this tick)")
        self._last_ok = result
        self._last_ok_ts = time.time()
        return result

    def _stale_or_empty(self, msg: str) -> Dict[str, Any]:
        """compute.py:424-446."""
        if self._last_ok is not None and (self._stale_ttl_s is None
                                          or (time.time() - self._last_ok_ts) <= self._stale_ttl_s):
            return {"metrics": self._last_ok["metrics"], "status_message": msg,
                    "rank_heatmap": self._last_ok["rank_heatmap"]}
        return _empty("No fresh step-combined data")

    # ------------------------------------------------------------------ core
    def _sid(self) -> int:
        if self.device.type != "cuda":
            return 0
        return int(torch.cuda.current_stream(self.device).cuda_stream)

    def _compute_impl(self, *, include_series: bool, include_rank_heatmap: bool) -> Dict[str, Any]:
        t0 = time.perf_counter()
        dev, sid, W = self.device, self._sid(), self.window_size
        lookback = max(W * self.lookback_factor, W)  # compute.py:366-368

        # ---- K7a per local rank + bounds exchange
        flat: List[float] = []
        for e in self.engines:
            i = e.combined_prepare(lookback, sid)
            flat += [int(i.n_rows), int(i.n_cand), int(i.lo), int(i.hi), int(i.latest_step)]
        infos: Dict[int, List[int]] = {}
        for p, row in enumerate(self.comm.all_gather_vec(flat, dev)):
            for l in range(self.L):
                infos[p * self.L + l] = [int(round(x)) for x in row[l * _INFO_LEN:(l + 1) * _INFO_LEN]]
        ranks = sorted(r for r, v in infos.items() if v[0] > 0)  # compute.py:139-143
        if not ranks:
            return _empty("No ranks available")
        completed = min(infos[r][4] for r in ranks)               # compute.py:147
        glo = max(infos[r][2] for r in ranks)
        ghi = min(min(infos[r][3] for r in ranks), completed)
        if ghi < glo:
            return _empty("No common step window yet")
        span = ghi - glo + 1

        # ---- K7b presence -> intersection -> K7c select + window sums
        presence = None
        for e in self.engines:
            p = torch.empty(span, dtype=torch.uint8, device=dev)
            e.combined_presence(glo, span, p, sid)
            presence = p if presence is None else torch.minimum(presence, p)
        self.comm.all_reduce_min_(presence)
        flat = []
        aligns = []
        for e in self.engines:
            a = e.combined_select(glo, span, presence, W, sid)
            aligns.append(a)
            flat += [int(a.n_common), int(a.n_rows)] + [float(x) for x in a.sums]
        n = int(aligns[0].n_common)
        if n == 0:
            return _empty("No common step window yet")
        sums: Dict[int, List[float]] = {}
        for p, row in enumerate(self.comm.all_gather_vec(flat, dev)):
            for l in range(self.L):
                v = row[l * _ALIGN_LEN:(l + 1) * _ALIGN_LEN]
                r = p * self.L + l
                if r in ranks and int(round(v[1])) > 0:
                    sums[r] = [float(x) for x in v[2:8]]
        present = [r for r in ranks if r in sums]
        coverage = {"expected_steps": W, "steps_used": n, "completed_step": int(completed),
                    "world_size": len(ranks), "ranks_present": len(present),
                    "incomplete": len(present) < len(ranks)}

        # ---- K7d series (CLI mode): one row exchange, every rank reduces the tiny window
        steps: List[int] = []
        series = None
        if include_series:
            series, steps = self._series(present, aligns, n, glo, presence, sid)

        result = self._assemble(present, sums, coverage, steps, series, include_rank_heatmap)
        self.last_timings_ms = {"tick": (time.perf_counter() - t0) * 1e3}
        return result

    def _series(self, present, aligns, n, glo, presence, sid):
        dev = self.device
        local = torch.zeros(self.L, n * 8, dtype=torch.float64, device=dev)
        owner = None
        for l, e in enumerate(self.engines):
            if int(aligns[l].n_rows) > 0:
                local[l].copy_(e.combined_rows_tensor(n))
                owner = e if owner is None else owner
        if self.comm.world == 1:
            allrows = local.view(1, self.L, n * 8)
        else:
            allrows = torch.empty(self.comm.world, self.L, n * 8, dtype=torch.float64, device=dev)
            self.comm.all_gather_into(allrows.view(-1), local.view(-1))
        ptrs = [allrows[r // self.L, r % self.L].data_ptr() for r in present]
        out = torch.empty(18, n, dtype=torch.float64, device=dev)
        self.engines[0].combined_series(ptrs, n, out, sid)
        series = out.cpu().numpy()
        if owner is not None:
            steps = owner.combined_steps(n, sid)
        else:  # this process holds no rows: the reduced presence bytes name the steps
            steps = (torch.nonzero(presence).flatten()[-n:] + glo).cpu().tolist()
        return series, [int(s) for s in steps]

    # ------------------------------------------------------------------ rank-level (O(R) scalars)
    def _assemble(self, present, sums, coverage, steps, series, include_rank_heatmap):
        col = lambda k, r: sums[r][_COL[k]]  # noqa: E731
        by_key: Dict[str, Dict[int, float]] = {k: {r: col(k, r) for r in present}
                                               for k in DEFAULT_METRIC_KEYS}
        # compute.py:177-187
        wait = {r: max(0.0, col("step_time", r) - col("h2d", r) - col("forward", r)
                       - col("backward", r) - col("optimizer_step", r)) for r in present}
        by_key["wait_proxy"] = wait
        # compute.py:189-201
        scores = {r: col("dataloader_fetch", r)
                  + max(col("step_time", r),
                        col("h2d", r) + col("forward", r) + col("backward", r) + col("optimizer_step", r))
                  for r in present}
        worst_rank = max(scores, key=scores.get) if scores else None
        median_rank = None
        if scores:  # compute.py:628-660
            target = float(np.median(np.array(list(scores.values()), dtype=np.float64)))
            median_rank = min(scores, key=lambda r: (abs(scores[r] - target), scores[r], r))

        metrics: Dict[str, Any] = {}
        for k in self.metric_keys:
            ser = None
            if series is not None:
                c = _COL[k]
                ser = {"steps": list(steps), "median": series[c * 3].tolist(),
                       "worst": series[c * 3 + 1].tolist(), "sum": series[c * 3 + 2].tolist()}
            m = self._metric(k, by_key[k], present, coverage, ser)
            if k == "step_time" and worst_rank is not None:
                m["summary"]["worst_rank"] = int(worst_rank)  # compute.py:225-232
            metrics[k] = m
        metrics["wait_proxy"] = self._metric("wait_proxy", wait, present, coverage, None)

        heat = None
        if include_rank_heatmap and metrics:  # compute.py:263-297
            keys = [k for k in self.heatmap_keys if k in by_key]
            rows = [{"rank": int(r), "sums_ms": {k: float(by_key[k].get(r, 0.0)) for k in keys}}
                    for r in present]
            rows.sort(key=lambda row: (scores.get(row["rank"], 0.0), row["sums_ms"].get("step_time", 0.0),
                                       row["sums_ms"].get("dataloader_fetch", 0.0)), reverse=True)
            heat = {"window_size": self.window_size, "steps_used": coverage["steps_used"],
                    "metric_keys": keys, "rows": rows,
                    "sort_by": ["overall_score", "step_time", "dataloader_fetch"]}
        status = "OK"
        if worst_rank is not None:
            status += f" | overall_worst_rank=r{worst_rank}"
        if median_rank is not None:
            status += f" | overall_median_rank=r{median_rank}"
        return {"metrics": list(metrics.values()), "status_message": status, "rank_heatmap": heat}

    @staticmethod
    def _metric(key, rank_sums, ranks, coverage, series):
        """compute.py:533-626."""
        arr = np.array([float(rank_sums.get(r, 0.0)) for r in ranks], dtype=np.float64)
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
        return {"metric": str(key), "clock": "mixed", "series": series,
                "summary": {"window_size": int(coverage["expected_steps"]),
                            "steps_used": int(coverage["steps_used"]),
                            "median_total": median_total, "worst_total": worst_total,
                            "worst_rank": worst_rank, "skew_ratio": float(skew_ratio),
                            "skew_pct": float(skew_pct)},
                "coverage": coverage}
