"""CPU, build container only: a database written by traceml_b200.compat.SQLiteCompatWriter
is consumed by the reference's OWN final-report sections and yields the golden payloads
(the kept consumers keep working on our output)."""
import os
import sys

import pytest

REF_SRC = "/root/reference/src"
if os.path.isdir(REF_SRC) and REF_SRC not in sys.path:
    sys.path.append(REF_SRC)
pytest.importorskip("traceml.reporting.sections.step_time", reason="reference not importable here")

from helpers import assert_struct, load_golden, plain, proc_replay_for, step_replay_for  # noqa: E402
from traceml_b200 import records as rec_mod  # noqa: E402
import replay  # noqa: E402
from traceml_b200.compat import SQLiteCompatWriter  # noqa: E402


def _ident(r, n):
    return {"global_rank": r, "local_rank": r, "node_rank": 0, "hostname": "b200-box",
            "local_world_size": n, "world_size": n}


@pytest.mark.parametrize("case", ["input_straggler_r4", "duplicates_r2", "ragged_r4_w64", "mem_creep_confirmed_r4"])
def test_reference_sections_read_our_database(tmp_path, case):
    from traceml.reporting.sections.step_memory import StepMemorySummarySection
    from traceml.reporting.sections.step_time import StepTimeSummarySection

    g = load_golden(case)
    recs = step_replay_for(g)
    db = str(tmp_path / "telemetry")
    for r in sorted(recs):
        w = SQLiteCompatWriter(db, _ident(r, g["ranks"]), pid=1000 + r)
        w.write_step_time([rec_mod.step_record_to_wire(x, device=f"cuda:{r}") for x in recs[r]])
        w.write_step_memory([rec_mod.step_record_to_memory_wire(x, device=f"cuda:{r}") for x in recs[r]])
        w.close()
    res = StepTimeSummarySection(max_rows=g["window"]).build(db)
    assert_struct(plain(res.payload), g["step_time"]["payload"], "step_time.payload", rel=0.0)
    assert res.text == g["step_time"]["text"]
    mres = StepMemorySummarySection(window_size=g["window"]).build(db)
    assert_struct(plain(mres.payload), g["step_memory"]["payload"], "step_memory.payload", rel=0.0)


def test_reference_process_section_reads_our_database(tmp_path):
    from traceml.reporting.sections.process import ProcessSummarySection

    g = load_golden("proc_overhang_r4")
    procs = proc_replay_for(g)
    db = str(tmp_path / "telemetry")
    for r in sorted(procs):
        w = SQLiteCompatWriter(db, _ident(r, g["ranks"]))
        w.write_process([rec_mod.proc_record_to_wire(x, pid=1000 + r, ram_total=replay.PROC_RAM_TOTAL_BYTES,
                                                     gpu_count=g["ranks"], device_index=r) for x in procs[r]])
        w.close()
    res = ProcessSummarySection(max_process_rows=g["max_rows"]).build(db)
    assert_struct(plain(res.payload), g["process"]["payload"], "process.payload", rel=0.0)
    assert res.text == g["process"]["text"]
