"""Sampler seam: reference-named samplers over one drain cursor -- table names,
bounded tables with an append counter, the sender's envelope, rows delivered exactly once."""
import numpy as np


class _Engine:
    """Stand-in for the drain half of the engine (records come from the replay generator)."""
    device = 0

    def __init__(self, recs, procs):
        self._recs, self._procs = recs, procs

    def drain(self):
        out, self._recs = self._recs, self._recs[:0]
        return out, 0

    def proc_drain(self):
        out, self._procs = self._procs, self._procs[:0]
        return out, 0


def test_samplers_publish_reference_tables_once():
    import replay
    from traceml_b200.samplers import (ProcessSampler, RecordTap, StepMemorySampler, StepTimeSampler,
                                       TableStore)

    recs = replay.make_step_replay("balanced", 1, 12, 3)[0]
    procs = replay.make_proc_replay("normal", 1, 5, 3)[0]
    tap = RecordTap(_Engine(recs, procs))
    st, sm = StepTimeSampler(tap), StepMemorySampler(tap)
    pr = ProcessSampler(tap, probe=type("P", (), {"sample": lambda self, e: None})())
    for s in (pr, st, sm):
        s.sample()
    p = st.collect_payload()
    assert p["sampler"] == "StepTimeSampler" and list(p["tables"]) == ["StepTimeTable"]
    assert {"rank", "global_rank", "local_rank", "world_size", "local_world_size", "node_rank",
            "hostname", "pid", "sampler", "timestamp", "tables"} == set(p)
    rows = p["tables"]["StepTimeTable"]
    assert [r["step"] for r in rows] == [int(x) for x in recs["step"]]
    assert set(rows[0]) == {"seq", "timestamp", "step", "events"}
    m = sm.collect_payload()["tables"]["step_memory"]
    assert [r["step"] for r in m] == [int(x) for x in recs["step"]]
    assert set(m[0]) == {"seq", "ts", "model_id", "device", "step", "peak_alloc", "peak_resv"}
    assert len(pr.collect_payload()["tables"]["ProcessTable"]) == 5
    # nothing new -> no payload; rows are never delivered twice
    for s in (pr, st, sm):
        s.sample()
        assert s.collect_payload() is None
    db = TableStore("x", max_rows=3)
    for i in range(5):
        db.add_record("t", i)
    assert list(db.all_tables()["t"]) == [2, 3, 4] and db.get_append_count("t") == 5
