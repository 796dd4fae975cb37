"""Oracle: trailing-window trend evidence.  TEST INFRASTRUCTURE ONLY.

Restates ``src/traceml/analytics/trends/core.py:38-146`` and the band layout
of ``src/traceml/analytics/trends/schema.py:27-77`` (warm-up trim, then the
baseline / mid / recent band means of the retained series).
"""

from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

# schema.py:27-62 -- (start_frac, end_frac) of the stable (post warm-up) series
BASELINE_BAND = (0.15, 0.25)
MID_BAND = (0.45, 0.55)
RECENT_BAND = (0.90, 1.00)


def band_bounds(n: int, band: Tuple[float, float]) -> Tuple[int, int]:
    """Index range [start, end) of one band in a series of n points.

    core.py:38-49: floor(n*start), ceil(n*end), clamped so the slice is
    non-empty and inside the series.
    """
    start = int(math.floor(n * float(band[0])))
    end = int(math.ceil(n * float(band[1])))
    start = max(0, min(start, n - 1))
    end = max(start + 1, min(end, n))
    return start, end


def _mean(values: Sequence[float]) -> float:
    # core.py:34-35 -- plain left-to-right Python sum
    return float(sum(values) / max(1, len(values)))


def trend_evidence(
    series: Sequence[float],
    *,
    min_points: int = 200,
    warmup_frac: float = 0.10,
    history_limit: Optional[int] = 10_000,
) -> Optional[Dict[str, object]]:
    """core.py:51-115.  Returns None when the series is too short."""
    finite = []
    for v in series:
        try:
            x = float(v)
        except Exception:
            continue
        if math.isfinite(x):
            finite.append(x)
    values = finite
    if len(values) < int(min_points):
        return None
    truncated = False
    if history_limit is not None and len(values) > int(history_limit):
        values = values[-int(history_limit):]
        truncated = True
    if len(values) < int(min_points):
        return None
    warm = int(math.floor(len(values) * float(warmup_frac)))
    stable = values[warm:] if warm > 0 else values
    if len(stable) < int(min_points):
        return None
    n = len(stable)
    b0, b1 = band_bounds(n, BASELINE_BAND)
    m0, m1 = band_bounds(n, MID_BAND)
    r0, r1 = band_bounds(n, RECENT_BAND)
    base = _mean(stable[b0:b1])
    mid = _mean(stable[m0:m1])
    recent = _mean(stable[r0:r1])
    d_base = recent - base
    d_mid = recent - mid
    return {
        "points_seen": len(finite),
        "points_used": len(values),
        "truncated": truncated,
        "baseline_avg": base,
        "mid_avg": mid,
        "recent_avg": recent,
        "delta_vs_baseline": d_base,
        "delta_vs_mid": d_mid,
        "delta_pct_vs_baseline": None if abs(base) <= 1e-12 else d_base / base,
        "delta_pct_vs_mid": None if abs(mid) <= 1e-12 else d_mid / mid,
    }


def trend_pct(series: Sequence[float], **kw) -> Optional[float]:
    """core.py:118-129."""
    ev = trend_evidence(series, **kw)
    return None if ev is None else ev["delta_pct_vs_baseline"]


def format_trend_pct(pct: Optional[float], *, deadband_pct: float = 0.02) -> str:
    """core.py:132-146."""
    if pct is None:
        return "—"
    if abs(pct) < float(deadband_pct):
        return f"~ {pct * 100:+.1f}%"
    return f"{'↑' if pct > 0 else '↓'} {pct * 100:+.1f}%"
