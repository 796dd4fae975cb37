"""Live tick (StepCombined twin) on CPU: the oracle against the golden vectors the
unmodified reference produced (tests/golden/live, make_live_golden.py), and the host
orchestration of ``traceml_b200.live`` over the numpy engine double -- single
process (several local ranks) and world_size 2 over gloo."""
import json
import os
import sys
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
LIVE = os.path.join(HERE, "golden", "live")
CASES = json.load(open(os.path.join(LIVE, "INDEX.json")))["cases"]
MEM_CASES = json.load(open(os.path.join(LIVE, "INDEX.json")))["mem_cases"]


def mem_rows(records):
    out = {}
    for r in records:
        has = (records[r]["flags"] & 1) != 0
        out[r] = [(int(s), (float(a) if h else None), (float(v) if h else None))
                  for s, a, v, h in zip(records[r]["step"], records[r]["peak_alloc"],
                                        records[r]["peak_resv"], has)]
    return out


def strip_dev(res):
    for m in res["metrics"]:
        m.pop("device", None)
    return res


def load(name):
    with open(os.path.join(LIVE, name + ".json")) as f:
        return json.load(f)


def wire_rows(records):
    from traceml_b200 import records as rec_mod
    return {r: [rec_mod.step_record_to_wire(x, device=f"cuda:{r}") for x in records[r]] for r in records}


@pytest.mark.parametrize("name", CASES)
def test_live_oracle_matches_reference_golden(name):
    from helpers import assert_struct, plain
    from oracle import live_oracle
    import replay

    g = load(name)
    recs = replay.make_step_replay(g["scenario"], g["ranks"], g["steps"], g["seed"])
    assert replay.replay_digest(recs) == g["digest"]
    rows = wire_rows(recs)
    cli = live_oracle.live_step_time(rows, window=g["window"], include_series=True)
    dash = live_oracle.live_step_time(rows, window=g["window"], include_series=False,
                                      include_rank_heatmap=True)
    assert_struct(plain(cli), g["cli"], name + ".cli", rel=0.0)
    assert_struct(plain(dash), g["dashboard"], name + ".dashboard", rel=0.0)


@pytest.mark.parametrize("name", CASES)
def test_live_host_logic_with_engine_double(name):
    """All ranks as local engines of one process: the host side of live.py (bounds,
    intersection, assembly) must reproduce the reference's result exactly."""
    from fake_engine import FakeEngine
    from helpers import assert_struct, plain
    import replay
    from traceml_b200.live import StepCombinedComputer

    g = load(name)
    recs = replay.make_step_replay(g["scenario"], g["ranks"], g["steps"], g["seed"])
    engines = [FakeEngine(recs[r]) for r in sorted(recs)]
    comp = StepCombinedComputer(engines, window_size=g["window"], device=torch.device("cpu"))
    cli = comp._compute_impl(include_series=True, include_rank_heatmap=False)
    dash = comp._compute_impl(include_series=False, include_rank_heatmap=True)
    assert_struct(plain(cli), g["cli"], name + ".cli", rel=0.0)
    assert_struct(plain(dash), g["dashboard"], name + ".dashboard", rel=0.0)


@pytest.mark.parametrize("name", MEM_CASES)
def test_live_memory_oracle_and_host_logic(name):
    """Step-memory panel: oracle == reference golden, and live.py over the engine double
    (all ranks local) == reference golden."""
    from fake_engine import FakeEngine
    from helpers import assert_struct, plain
    from oracle import live_oracle
    import replay
    from traceml_b200.live import StepMemoryCombinedComputer

    g = load(name)
    recs = replay.make_step_replay(g["scenario"], g["ranks"], g["steps"], g["seed"])
    assert replay.replay_digest(recs) == g["digest"]
    o = live_oracle.live_step_memory(mem_rows(recs), window=g["window"], gpu_available=g["gpu_available"])
    assert_struct(plain(o), g["result"], name + ".oracle", rel=0.0)
    engines = [FakeEngine(recs[r]) for r in sorted(recs)]
    comp = StepMemoryCombinedComputer(engines, window_size=g["window"], gpu_available=g["gpu_available"],
                                      device=torch.device("cpu"))
    got = strip_dev(comp._compute_impl())
    assert_struct(plain(got), g["result"], name + ".host", rel=0.0)


def test_live_memory_far_ahead_rank_widens_lookback():
    """One rank thousands of steps ahead: its in-range rows lie beyond the first look-back."""
    from fake_engine import FakeEngine
    from helpers import assert_struct, plain
    from oracle import live_oracle
    import replay
    from traceml_b200.live import StepMemoryCombinedComputer

    recs = replay.make_step_replay("balanced", 2, 6000, 3)
    recs[1] = recs[1][:900]  # rank 1 lags: completed = 900, window 801..900 far behind rank 0's tail
    engines = [FakeEngine(recs[r]) for r in sorted(recs)]
    comp = StepMemoryCombinedComputer(engines, window_size=100, gpu_available=True, device=torch.device("cpu"))
    got = strip_dev(comp._compute_impl())
    ref = live_oracle.live_step_memory(mem_rows(recs), window=100, gpu_available=True)
    assert ref["metrics"] and ref["metrics"][0]["coverage"]["ranks_present"] == 2
    assert_struct(plain(got), plain(ref), "widen", rel=0.0)


def test_live_stale_handling():
    """compute.py:103-123, 424-446: an empty tick serves the last good result."""
    from fake_engine import FakeEngine
    import replay
    from traceml_b200.live import StepCombinedComputer

    recs = replay.make_step_replay("balanced", 2, 60, 1)
    engines = [FakeEngine(recs[r]) for r in sorted(recs)]
    comp = StepCombinedComputer(engines, window_size=20, device=torch.device("cpu"))
    assert comp.compute_cli()["status_message"].startswith("OK")
    good = comp.compute_cli()
    for e in engines:
        e.records = e.records[:0]
    stale = comp.compute_cli()
    assert stale["status_message"] == "STALE (no metrics this tick)"
    assert stale["metrics"] == good["metrics"]
    comp._stale_ttl_s = 0.0
    comp._last_ok_ts -= 1.0
    assert comp.compute_cli() == {"metrics": [], "status_message": "No fresh step-combined data",
                                  "rank_heatmap": None}
    with pytest.raises(ValueError):
        StepCombinedComputer(engines, metric_keys=["nope"], device=torch.device("cpu"))


def _worker(rank, world, name, init_file, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    from fake_engine import FakeEngine
    import replay
    from traceml_b200.live import StepCombinedComputer
    from traceml_b200.reduce import TorchDistComm

    g = load(name)
    recs = replay.make_step_replay(g["scenario"], g["ranks"], g["steps"], g["seed"])
    L = g["ranks"] // world
    engines = [FakeEngine(recs[rank * L + l]) for l in range(L)]
    comp = StepCombinedComputer(engines, TorchDistComm(), window_size=g["window"],
                                device=torch.device("cpu"))
    out = {"cli": comp._compute_impl(include_series=True, include_rank_heatmap=False),
           "dashboard": comp._compute_impl(include_series=False, include_rank_heatmap=True)}
    torch.save(out, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def _mem_worker(rank, world, name, init_file, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    from fake_engine import FakeEngine
    import replay
    from traceml_b200.live import StepMemoryCombinedComputer
    from traceml_b200.reduce import TorchDistComm

    g = load(name)
    recs = replay.make_step_replay(g["scenario"], g["ranks"], g["steps"], g["seed"])
    L = g["ranks"] // world
    engines = [FakeEngine(recs[rank * L + l]) for l in range(L)]
    comp = StepMemoryCombinedComputer(engines, TorchDistComm(), window_size=g["window"],
                                      gpu_available=g["gpu_available"], device=torch.device("cpu"))
    torch.save(strip_dev(comp._compute_impl()), os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["livemem_ragged_r4", "livemem_no_overlap_r2", "livemem_default_r8"])
def test_live_memory_two_process_gloo(name):
    from helpers import assert_struct, plain

    g = load(name)
    world = 2
    with tempfile.TemporaryDirectory() as td:
        init_file = os.path.join(td, "init")
        mp.spawn(_mem_worker, args=(world, name, init_file, td), nprocs=world, join=True)
        got = [torch.load(os.path.join(td, f"r{r}.pt"), weights_only=False) for r in range(world)]
    for r in range(world):
        assert_struct(plain(got[r]), g["result"], f"{name}.r{r}", rel=0.0)


@pytest.mark.parametrize("name", ["live_ragged_r4", "live_duplicates_r2", "live_wait_heavy_r8"])
def test_live_two_process_gloo(name):
    from helpers import assert_struct, plain

    g = load(name)
    world = 2
    with tempfile.TemporaryDirectory() as td:
        init_file = os.path.join(td, "init")
        mp.spawn(_worker, args=(world, name, init_file, td), nprocs=world, join=True)
        got = [torch.load(os.path.join(td, f"r{r}.pt"), weights_only=False) for r in range(world)]
    for r in range(world):
        assert_struct(plain(got[r]["cli"]), g["cli"], f"{name}.r{r}.cli", rel=0.0)
        assert_struct(plain(got[r]["dashboard"]), g["dashboard"], f"{name}.r{r}.dashboard", rel=0.0)


def _live_fuzz(n, seed):
    import random

    rng = random.Random(seed)
    sc = ["balanced", "input_straggler", "straggler", "wait_heavy", "ragged", "duplicates", "empty_rank",
          "no_overlap", "warmup", "mem_imbalance", "mem_creep_confirmed", "trend_worsening"]
    return [(rng.choice(sc), rng.choice([1, 2, 3, 5, 8]), rng.choice([30, 90, 260, 700, 2600]),
             rng.randrange(10_000), rng.choice([7, 50, 100, 400])) for _ in range(n)]


@pytest.mark.parametrize("scenario,R,S,seed,W", _live_fuzz(30, 515))
def test_live_host_logic_random(scenario, R, S, seed, W):
    """Seeded random cases: both live views over the engine double against their oracle."""
    from fake_engine import FakeEngine
    from helpers import assert_struct, plain
    from oracle import live_oracle
    import replay
    from traceml_b200.live import StepCombinedComputer, StepMemoryCombinedComputer

    recs = replay.make_step_replay(scenario, R, S, seed)
    engines = [FakeEngine(recs[r]) for r in sorted(recs)]
    rows = wire_rows(recs)
    t = StepCombinedComputer(engines, window_size=W, device=torch.device("cpu"))
    assert_struct(plain(t._compute_impl(include_series=True, include_rank_heatmap=False)),
                  plain(live_oracle.live_step_time(rows, window=W)), "time.cli", rel=0.0)
    assert_struct(plain(t._compute_impl(include_series=False, include_rank_heatmap=True)),
                  plain(live_oracle.live_step_time(rows, window=W, include_series=False,
                                                   include_rank_heatmap=True)), "time.dash", rel=0.0)
    m = StepMemoryCombinedComputer(engines, window_size=W, gpu_available=True, device=torch.device("cpu"))
    assert_struct(plain(strip_dev(m._compute_impl())),
                  plain(live_oracle.live_step_memory(mem_rows(recs), window=W, gpu_available=True)),
                  "mem", rel=0.0)
