"""``final_summary()`` -- in-process end-of-run summary.

The reference answers ``traceml.final_summary()`` with a file request/response
round trip to the aggregator process, which runs the sections over SQLite
(``src/traceml/sdk/summary_client.py:35``, ``aggregator/summary_service.py:27-114``).
Here every rank already holds its window in HBM, so the summary is one
collective call: all ranks enter, rank 0 (or every rank) gets the result.
"""
from __future__ import annotations

import json
import os
from typing import Any, Dict, Optional

from .runtime import disabled, get_engine, summary_window_rows


def _replace_with(path: str, text: str) -> None:
    """Readers never see a half-written artifact: sibling temp file, fsync, rename."""
    folder = os.path.dirname(os.path.abspath(path))
    os.makedirs(folder, exist_ok=True)
    tmp = os.path.join(folder, f".{os.path.basename(path)}.{os.getpid()}.tmp")
    try:
        with open(tmp, "w", encoding="utf-8") as fh:
            fh.write(text)
            fh.flush()
            os.fsync(fh.fileno())
        os.replace(tmp, path)
    finally:
        if os.path.exists(tmp):
            os.unlink(tmp)


def write_summary_artifacts(summary: Dict[str, Any], session_root: str) -> Dict[str, str]:
    """``<session>/final_summary.json`` (``json.dumps(indent=2)``) and ``final_summary.txt``:
    the artifact names and format of ``sdk/protocol.py:160-171`` / ``utils/atomic_io.py:18-63``
    that ``traceml compare`` and the launcher's end-of-run printout read."""
    root = os.path.abspath(session_root)
    paths = {"json": os.path.join(root, "final_summary.json"), "txt": os.path.join(root, "final_summary.txt")}
    _replace_with(paths["json"], json.dumps(summary, indent=2))
    _replace_with(paths["txt"], str(summary.get("text", "")))
    return paths


def final_summary(*, timeout_sec: float = 30.0, poll_interval_sec: float = 0.1,
                  print_text: bool = False, rank0_only: bool = True,
                  window_rows: Optional[int] = None,
                  session_root: Optional[str] = None) -> Optional[Dict[str, Any]]:
    """Collective over the default process group when one is initialised.  ``session_root``
    (default: ``$TRACEML_SESSION_ROOT`` if set): where rank 0 also writes the artifacts."""
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
    root = session_root or os.environ.get("TRACEML_SESSION_ROOT")
    if root and comm.index == 0:
        try:
            write_summary_artifacts(out, root)
        except OSError as exc:  # artifacts are best effort; the caller still gets the summary
            import sys

            print(f"[TraceML] could not write final summary artifacts: {exc}", file=sys.stderr)
    if print_text and out.get("text"):
        print(out["text"])
    return out


__all__ = ["final_summary", "write_summary_artifacts"]
