"""``final_summary()`` -- in-process end-of-run summary.

The reference answers ``traceml.final_summary()`` with a file request/response
round trip to the aggregator process, which runs the sections over SQLite
(``src/traceml/sdk/summary_client.py:35``, ``aggregator/summary_service.py:27-114``).
Here every rank already holds its window in HBM, so the summary is one
collective call: all ranks enter, rank 0 (or every rank) gets the result.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

from .runtime import disabled, get_engine, summary_window_rows


def final_summary(*, timeout_sec: float = 30.0, poll_interval_sec: float = 0.1,
                  print_text: bool = False, rank0_only: bool = True,
                  window_rows: Optional[int] = None) -> Optional[Dict[str, Any]]:
    """Collective over the default process group when one is initialised."""
    if disabled():
        return None
    import torch
    import torch.distributed as dist

    from .reduce import LocalComm, TorchDistComm
    from .reporting import build_final_summary
    from .sections import SummaryEngine

    eng = get_engine()
    torch.cuda.current_stream(torch.device("cuda", eng.device)).synchronize()
    comm = TorchDistComm() if (dist.is_available() and dist.is_initialized()
                               and dist.get_world_size() > 1) else LocalComm()
    res = SummaryEngine([eng], comm).build(window_rows or summary_window_rows(),
                                           window_rows or summary_window_rows())
    if rank0_only and comm.index != 0:
        return None
    out = build_final_summary(res)
    if print_text and out.get("text"):
        print(out["text"])
    return out


__all__ = ["final_summary"]
