"""GPU parity of the live tick (csrc/tml_combined.cuh + traceml_b200/live.py) through
the C-ABI, against the reference's own StepCombinedComputer outputs
(tests/golden/live) and the oracle.  Step ids, rank ids and status labels exact;
floats rel <= 1e-9 (in practice bit-equal: the sums run in the reference's order)."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import REL_TOL, assert_struct, plain

pytestmark = pytest.mark.gpu

LIVE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "live")
CASES = json.load(open(os.path.join(LIVE, "INDEX.json")))["cases"]


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return torch.device("cuda", 0)


def _engines(records, ring_slots=None):
    from traceml_b200.engine import Engine

    R = len(records)
    out = []
    for r in range(R):
        n = len(records[r])
        e = Engine(device=0, rank=r, world=R, ring_slots=ring_slots or max(64, n + 16), proc_slots=64)
        if n:
            e.load_steps(records[r])
        out.append(e)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("name", CASES)
def test_live_vs_reference_golden(cuda, name):
    import replay
    from traceml_b200.live import StepCombinedComputer

    g = json.load(open(os.path.join(LIVE, name + ".json")))
    recs = replay.make_step_replay(g["scenario"], g["ranks"], g["steps"], g["seed"])
    assert replay.replay_digest(recs) == g["digest"]
    engines = _engines(recs)
    try:
        comp = StepCombinedComputer(engines, window_size=g["window"])
        with torch.cuda.stream(comp._stream):
            cli = comp._compute_impl(include_series=True, include_rank_heatmap=False)
            dash = comp._compute_impl(include_series=False, include_rank_heatmap=True)
    finally:
        for e in engines:
            e.close()
    assert cli["status_message"] == g["cli"]["status_message"]
    assert_struct(plain(cli), g["cli"], name + ".cli", rel=REL_TOL)
    assert_struct(plain(dash), g["dashboard"], name + ".dashboard", rel=REL_TOL)
    # integer / label side and the rank order of the heat map: exact
    for a, b in zip(cli["metrics"], g["cli"]["metrics"]):
        assert a["summary"]["worst_rank"] == b["summary"]["worst_rank"]
        if b["series"]:
            assert a["series"]["steps"] == b["series"]["steps"]
    if g["dashboard"]["rank_heatmap"]:
        assert [r["rank"] for r in dash["rank_heatmap"]["rows"]] == \
               [r["rank"] for r in g["dashboard"]["rank_heatmap"]["rows"]]


def test_live_ring_wrap_and_oracle(cuda):
    """The ring wrapped many times; the tick must see only the newest look-back rows."""
    from oracle import live_oracle
    from traceml_b200 import records as rec_mod
    import replay
    from traceml_b200.live import StepCombinedComputer

    R, S, W = 4, 3000, 100
    recs = replay.make_step_replay("ragged", R, S, 5)
    engines = _engines(recs, ring_slots=1024)
    try:
        got = StepCombinedComputer(engines, window_size=W).compute_cli()
    finally:
        for e in engines:
            e.close()
    rows = {r: [rec_mod.step_record_to_wire(x, device=f"cuda:{r}") for x in recs[r]] for r in recs}
    ref = live_oracle.live_step_time(rows, window=W)
    assert_struct(plain(got), plain(ref), "wrap", rel=REL_TOL)


def test_live_tick_during_training(cuda):
    """A tick on its side stream while the step path keeps committing: it neither
    waits for the training stream nor perturbs the committed records."""
    import traceml_b200 as traceml
    from traceml_b200 import runtime
    from traceml_b200.live import StepCombinedComputer

    from traceml_b200.runtime import reset_trace_session_state

    reset_trace_session_state(0)
    traceml.init(mode="auto")
    eng = runtime.get_engine()
    torch.cuda.synchronize()
    eng.reset()
    model = torch.nn.Sequential(torch.nn.Linear(256, 256), torch.nn.ReLU(), torch.nn.Linear(256, 8)).cuda()
    opt = torch.optim.SGD(model.parameters(), lr=1e-3)
    comp = StepCombinedComputer([eng], window_size=16)
    x = torch.randn(64, 256, device="cuda")
    ticks = []
    for i in range(60):
        with traceml.trace_step(model):
            loss = model(x).square().mean()
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
        if i % 10 == 9:
            ticks.append(comp.compute_cli())
    torch.cuda.synchronize()
    final = comp.compute_cli()
    assert all(t["status_message"].startswith("OK") for t in ticks)
    used = [t["metrics"][0]["summary"]["steps_used"] for t in ticks]
    assert used[-1] == 16 and all(u <= 16 for u in used)
    st = {m["metric"]: m for m in final["metrics"]}
    assert st["step_time"]["series"]["steps"] == list(range(45, 61))
    assert st["step_time"]["summary"]["median_total"] > 0.0
    assert st["forward"]["summary"]["median_total"] > 0.0
    assert np.isclose(sum(st["step_time"]["series"]["sum"]), st["step_time"]["summary"]["worst_total"],
                      rtol=1e-12)


MEM_CASES = json.load(open(os.path.join(LIVE, "INDEX.json")))["mem_cases"]


def _strip_dev(res):
    for m in res["metrics"]:
        m.pop("device", None)
    return res


@pytest.mark.parametrize("name", MEM_CASES)
def test_live_memory_vs_reference_golden(cuda, name):
    import replay
    from traceml_b200.live import StepMemoryCombinedComputer

    g = json.load(open(os.path.join(LIVE, name + ".json")))
    recs = replay.make_step_replay(g["scenario"], g["ranks"], g["steps"], g["seed"])
    assert replay.replay_digest(recs) == g["digest"]
    engines = _engines(recs)
    try:
        comp = StepMemoryCombinedComputer(engines, window_size=g["window"], gpu_available=g["gpu_available"])
        with torch.cuda.stream(comp._stream):
            got = _strip_dev(comp._compute_impl())
    finally:
        for e in engines:
            e.close()
    assert got["status_message"] == g["result"]["status_message"]
    assert_struct(plain(got), g["result"], name, rel=REL_TOL)
    for a, b in zip(got["metrics"], g["result"]["metrics"]):
        assert a["summary"]["worst_rank"] == b["summary"]["worst_rank"]
        assert a["series"]["steps"] == b["series"]["steps"]
        assert a["series"]["median"] == b["series"]["median"]  # bytes: exact
        assert a["series"]["worst"] == b["series"]["worst"]


def test_live_memory_far_ahead_rank(cuda):
    """A rank thousands of steps ahead of the slowest: its in-range rows lie beyond the first
    look-back and the ring has wrapped; the tick widens locally and stays exact."""
    from oracle import live_oracle
    import replay
    from traceml_b200.live import StepMemoryMetricsComputer

    recs = replay.make_step_replay("balanced", 2, 6000, 3)
    recs[1] = recs[1][:900]
    engines = _engines(recs, ring_slots=8192)
    try:
        got = _strip_dev(StepMemoryMetricsComputer(engines, cli_window_size=100).compute_cli())
    finally:
        for e in engines:
            e.close()
    rows = {}
    for r in recs:
        rows[r] = [(int(s), float(a), float(v)) for s, a, v in
                   zip(recs[r]["step"], recs[r]["peak_alloc"], recs[r]["peak_resv"])]
    ref = live_oracle.live_step_memory(rows, window=100, gpu_available=True)
    assert ref["metrics"][0]["coverage"]["ranks_present"] == 2
    assert_struct(plain(got), plain(ref), "far_ahead", rel=REL_TOL)


def _live_fuzz(n, seed):
    import random

    rng = random.Random(seed)
    sc = ["balanced", "input_straggler", "straggler", "wait_heavy", "ragged", "duplicates", "empty_rank",
          "no_overlap", "warmup", "mem_imbalance", "mem_creep_confirmed", "trend_worsening"]
    return [(rng.choice(sc), rng.choice([1, 2, 3, 5, 8]), rng.choice([30, 90, 260, 700, 2600]),
             rng.randrange(10_000), rng.choice([7, 50, 100, 400]), rng.choice([None, 512]))
            for _ in range(n)]


@pytest.mark.parametrize("scenario,R,S,seed,W,slots", _live_fuzz(16, 616))
def test_live_views_random_vs_oracle(cuda, scenario, R, S, seed, W, slots):
    """Seeded random cases through the K7 kernels (ring wrap included) against the live oracle."""
    from oracle import live_oracle
    from traceml_b200 import records as rec_mod
    import replay
    from traceml_b200.live import StepCombinedComputer, StepMemoryCombinedComputer

    recs = replay.make_step_replay(scenario, R, S, seed)
    engines = _engines(recs, ring_slots=slots)
    try:
        cli = StepCombinedComputer(engines, window_size=W).compute_cli()
        t = StepCombinedComputer(engines, window_size=W)
        with torch.cuda.stream(t._stream):
            dash = t._compute_impl(include_series=False, include_rank_heatmap=True)
        m = StepMemoryCombinedComputer(engines, window_size=W, gpu_available=True)
        with torch.cuda.stream(m._stream):
            mem = _strip_dev(m._compute_impl())
    finally:
        for e in engines:
            e.close()
    # the look-back never reaches past what the ring retains (minus the in-flight guard)
    kept = {r: (recs[r][-(slots - 8):] if slots and len(recs[r]) > slots - 8 else recs[r]) for r in recs}
    rows = {r: [rec_mod.step_record_to_wire(x, device=f"cuda:{r}") for x in kept[r]] for r in kept}
    ref = live_oracle.live_step_time(rows, window=W)
    if ref["metrics"]:
        assert_struct(plain(cli), plain(ref), "cli", rel=REL_TOL)
    else:  # compute_cli() wraps an empty tick with its stale message
        assert cli["metrics"] == [] and cli["status_message"] == "No fresh step-combined data"
    assert_struct(plain(dash), plain(live_oracle.live_step_time(rows, window=W, include_series=False,
                                                                include_rank_heatmap=True)), "dash", rel=REL_TOL)
    mrows = {r: [(int(s), float(a), float(v)) if (f & 1) else (int(s), None, None)
                 for s, a, v, f in zip(kept[r]["step"], kept[r]["peak_alloc"], kept[r]["peak_resv"],
                                       kept[r]["flags"])] for r in kept}
    assert_struct(plain(mem), plain(live_oracle.live_step_memory(mrows, window=W, gpu_available=True)),
                  "mem", rel=REL_TOL)
