"""tml_sections_json (csrc/tml_sections.cpp, host C++: no GPU needed) against
traceml_b200/sections.py on every golden case: the native section objects must equal the
Python-assembled ones exactly, and both equal the reference's golden outputs."""
import json
import os

import pytest
import torch

from helpers import assert_struct, golden_cases, plain, proc_replay_for, step_replay_for

STEP = golden_cases("step")
PROC = golden_cases("process")


def fill_run_out(red, window, proc_rows):
    """ReduceOutput (Python driver) -> the struct tml_reduce_run would have produced."""
    from traceml_b200 import _abi

    o = _abi.ReduceRunOut()
    R = len(red.ranks)
    o.n_ranks = R
    for r in red.ranks:
        d, i = red.infos[r], o.infos[r]
        i.n_retained, i.latest_step, i.monotone, i.dup_rows = d["n_retained"], d["latest_step"], d["monotone"], d["dup_rows"]
        for q in range(2):
            i.n_rows[q], i.n_cand[q], i.lo[q], i.hi[q] = d["n_rows"][q], d["n_cand"][q], d["lo"][q], d["hi"][q]
            i.dense[q] = d["dense"][q]
        for q in range(7):
            i.t_sums[q] = d["t_sums"][q]
        i.t_count, i.n_both = d["t_count"], d["n_both"]
        if proc_rows and r in red.proc_aggs:
            for f, _ in _abi.ProcAgg._fields_:
                setattr(o.procs[r], f, red.proc_aggs[r][f])
    for k, res in ((o.time, red.time), (o.mem, red.mem)):
        k.observed, k.n_common = res.observed, res.n_common
        k.n_used = len(res.used)
        k.start_step, k.end_step = res.start_step or 0, res.end_step or 0
        for idx, r in enumerate(res.used):
            w = res.windows[r]
            k.used[idx], k.n_rows[idx] = r, w.n_rows
            for q in range(7):
                k.t_sums[idx][q] = w.t_sums[q]
            for q in range(4):
                k.m_sums[idx][q] = w.m_sums[q]
        if res.band_sum is not None:
            k.has_bands = 1
            for s in range(16):
                for b in range(3):
                    k.band_sum[s][b], k.band_cnt[s][b] = res.band_sum[s][b], res.band_cnt[s][b]
                k.tail_first[s], k.tail_last[s] = res.tail_first[s], res.tail_last[s]
    return o


def both(records, procs, window):
    from fake_engine import FakeEngine
    import replay
    from traceml_b200 import _abi, sections

    R = len(records) if records is not None else len(procs)
    empty = replay.make_step_replay("balanced", 1, 0, 0)[0]
    engines = [FakeEngine(records[r] if records is not None else empty,
                          procs[r] if procs is not None else None) for r in range(R)]
    se = sections.SummaryEngine(engines, ram_total=replay.PROC_RAM_TOTAL_BYTES, gpu_count=R)
    se.reducer.device = torch.device("cpu")
    py = se.build(window, window)
    red = py.pop("reduce")
    nat = _abi.sections_json(fill_run_out(red, window, window), replay.PROC_RAM_TOTAL_BYTES, R, window, window)
    return py, nat


@pytest.mark.parametrize("g", STEP, ids=[g["case"] for g in STEP])
def test_native_sections_equal_python_sections(g):
    recs = step_replay_for(g)
    py, nat = both(recs, None, g["window"])
    for sec in ("step_time", "step_memory"):
        assert_struct(plain(nat[sec]), plain(py[sec]), sec, rel=0.0)
    # and the golden: the reference's own data / diagnosis
    assert_struct(plain(nat["step_time"]["data"]), g["step_time"]["data"], "golden.data")
    assert_struct(plain(nat["step_time"]["diagnosis"]), g["step_time"]["diagnosis"], "golden.diagnosis")
    for k in ("average", "median", "worst"):
        assert_struct(plain(nat["step_time"]["global"][k]), g["step_time"]["payload"]["global"][k], f"golden.global.{k}")
    assert isinstance(next(iter(nat["step_time"]["data"]["aligned_summary"]), 0), int)


@pytest.mark.parametrize("g", PROC, ids=[g["case"] for g in PROC])
def test_native_process_section_equals_python(g):
    procs = proc_replay_for(g)
    py, nat = both(None, procs, g["max_rows"])
    assert_struct(plain(nat["process"]), plain(py["process"]), "process", rel=0.0)
    assert_struct(plain(nat["process"]["primary"]), g["process"]["diagnosis"]["primary"], "golden.primary")
    assert_struct(plain(nat["process"]["issues"]), g["process"]["diagnosis"]["issues"], "golden.issues")


@pytest.mark.parametrize("pattern", [(0, 0, 1, 1), (0, 0, 0, 0, 0, 0), (0, 1, 0, 1, 0), (1, 0), (0, 1, 2, 0, 1, 2, 0)])
def test_tie_breaks_with_identical_ranks(pattern):
    """Ranks holding identical records: every per-rank value ties, so median / worst ranks are
    decided purely by the tie-break rules (|delta|, value, rank / lowest rank wins) -- the
    native rollups must make the same choices as sections.py."""
    import replay

    base = replay.make_step_replay("input_straggler", 3, 260, seed=5)
    recs = {r: base[p].copy() for r, p in enumerate(pattern)}
    pbase = replay.make_proc_replay("imbalance", 3, 120, seed=5)
    procs = {r: pbase[p].copy() for r, p in enumerate(pattern)}
    py, nat = both(recs, procs, 10_000)
    for sec in ("step_time", "step_memory", "process"):
        assert_struct(plain(nat[sec]), plain(py[sec]), sec, rel=0.0)
    # and against the oracle (the reference's arithmetic) for the public rollup
    from helpers import oracle_time_rows
    from oracle import step_time_oracle

    o = step_time_oracle.step_time_section(oracle_time_rows(recs, 10_000), max_rows=10_000)
    for k in ("average", "median", "worst"):
        assert_struct(plain(nat["step_time"]["global"][k]), plain(o["global"][k]), f"oracle.global.{k}", rel=0.0)
    assert_struct(plain(nat["step_time"]["diagnosis"]), plain(o["diagnosis"]), "oracle.diagnosis", rel=0.0)


def test_sections_json_from_several_threads():
    """The JSON strings of one call come from a per-thread bump arena (tml_internal.h) that is
    rewound when the outermost entry point returns: calls racing on different threads (ctypes
    drops the GIL) must not see each other's memory."""
    import ctypes as C
    import threading

    from fake_engine import FakeEngine
    import replay
    from traceml_b200 import _abi, sections

    outs = {}
    for R in (1, 3, 8):
        W = 300
        recs = replay.make_step_replay("input_straggler", R, W, seed=R)
        procs = replay.make_proc_replay("normal", R, W, seed=R)
        se = sections.SummaryEngine([FakeEngine(recs[r], procs[r]) for r in range(R)],
                                    ram_total=replay.PROC_RAM_TOTAL_BYTES, gpu_count=R)
        se.reducer.device = torch.device("cpu")
        red = se.build(W, W).pop("reduce")
        outs[R] = (fill_run_out(red, W, W), R, W)
    lib = _abi.lib()
    want = {}
    for R, (o, _, W) in outs.items():
        buf = C.create_string_buffer(1 << 18)
        args = _abi.SectionsArgs(float(replay.PROC_RAM_TOTAL_BYTES), R, W, W, 0)
        assert lib.tml_sections_json(C.byref(o), C.byref(args), buf, len(buf)) == 0
        want[R] = buf.value
    bad = []

    def worker(R):
        o, _, W = outs[R]
        buf = C.create_string_buffer(1 << 18)
        args = _abi.SectionsArgs(float(replay.PROC_RAM_TOTAL_BYTES), R, W, W, 0)
        for _ in range(300):
            if lib.tml_sections_json(C.byref(o), C.byref(args), buf, len(buf)) != 0 or buf.value != want[R]:
                bad.append(R)
                return

    threads = [threading.Thread(target=worker, args=(R,)) for R in (1, 3, 8, 1, 3, 8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not bad, bad
