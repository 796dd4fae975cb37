"""CPU: the timer-path oracle reproduces the reference's own wire rows
(tests/golden/timer_rows.json, written by tests/golden/make_timer_golden.py from
the unmodified reference under a deterministic clock)."""
import json
import os
import time

import torch


def test_timer_oracle_matches_reference_rows(monkeypatch):
    import importlib.util
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("mtg", os.path.join(here, "golden", "make_timer_golden.py"))
    saved = list(sys.path)
    mtg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mtg)
    sys.path[:] = [p for p in saved]
    got_t, got_m = mtg.run_oracle()
    with open(os.path.join(here, "golden", "timer_rows.json")) as fh:
        ref = json.load(fh)
    assert [r["step"] for r in got_t] == ref["steps"] == [1, 2, 3, 4, 5, 5, 6]
    for g, r in zip(mtg.norm(got_t), ref["step_time"]):
        assert set(g["events"]) == set(r["events"])
        for name, by_dev in r["events"].items():
            for dev, e in by_dev.items():
                assert g["events"][name][dev]["n_calls"] == e["n_calls"]
                assert abs(g["events"][name][dev]["duration_ms"] - e["duration_ms"]) < 1e-6
    assert mtg.norm(got_m) == ref["step_memory"]
