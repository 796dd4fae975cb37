"""The large-window (numpy) oracle is pinned bit-for-bit to the row-level oracle, which is pinned
bit-for-bit to the unmodified reference (tests/golden/make_golden.py).  ``==`` on every float."""
import numpy as np
import pytest

import replay
from helpers import golden_cases, oracle_mem_rows, oracle_time_rows, plain, step_replay_for, strip_device
from oracle import fast_oracle, step_memory_oracle, step_time_oracle

STEP = golden_cases("step")


def assert_same(a, b, path=""):
    if isinstance(a, dict) and isinstance(b, dict):
        assert set(a) == set(b), f"{path}: keys {sorted(set(a) ^ set(b))}"
        for k in a:
            assert_same(a[k], b[k], f"{path}.{k}")
    elif isinstance(a, (list, tuple)) and isinstance(b, (list, tuple)):
        assert len(a) == len(b), f"{path}: len {len(a)} != {len(b)}"
        for i, (x, y) in enumerate(zip(a, b)):
            assert_same(x, y, f"{path}[{i}]")
    else:
        assert a == b or (a != a and b != b), f"{path}: {a!r} != {b!r}"


def compare(records, window):
    live = {r: v for r, v in records.items()}
    try:
        ft = fast_oracle.step_time_section(live, max_rows=window)
        fm = fast_oracle.step_memory_section(live, window_size=window)
    except NotImplementedError:
        pytest.skip("duplicated step ids: row-level oracle only")
    st = step_time_oracle.step_time_section(oracle_time_rows(records, window), max_rows=window)
    sm = step_memory_oracle.step_memory_section(oracle_mem_rows(records), window_size=window)
    d_f, d_s = ft["data"], st["data"]
    for k in ("training_steps", "latest_step_observed", "aligned_summary", "aligned_window",
              "per_global_rank_summary", "max_rows"):
        assert_same(plain(d_f[k]), plain(d_s[k]), f"time.data.{k}")
    for k in ("diagnosis", "global", "overview"):
        assert_same(plain(ft[k]), plain(st[k]), f"time.{k}")
    # element-wise series: the row-level oracle keeps them inside diagnose_summary's metrics
    if st["data"]["aligned_summary"]:
        full = step_time_oracle.diagnose_summary(
            step_time_oracle.rank_signals_from_summary(st["data"]["aligned_summary"]), max_rows=st["data"]["max_rows"],
            per_rank_step_metrics=st["data"]["aligned_step_metrics"], return_metrics=True)
        for m, key in enumerate(fast_oracle._TIME_KEYS):
            ms = [x for x in full.get("_metrics", []) if x["metric"] == key]
            if ms and ms[0]["series"]:
                assert ms[0]["series"]["median"] == ft["_series"][2 * m].tolist(), key
                assert ms[0]["series"]["worst"] == ft["_series"][2 * m + 1].tolist(), key
                assert ms[0]["series"]["steps"] == [int(s) for s in ft["_steps"]]
    assert fm["training_steps"] == sm["training_steps"]
    assert fm["window"]["n_steps"] == len(sm["window"]["steps"])
    assert fm["window"]["global_ranks_seen"] == sm["window"]["global_ranks_seen"]
    assert fm["window"]["global_ranks_used"] == sm["window"]["global_ranks_used"]
    assert_same(plain(fm["per_global_rank"]), plain(sm["per_global_rank"]), "mem.means")
    assert_same(plain(strip_device(fm["diagnosis"])), plain(strip_device(sm["diagnosis"])), "mem.diagnosis")
    assert_same(plain(fm["global"]), plain(sm["global"]), "mem.global")
    for a, b in zip(fm["metrics"], sm["metrics"]):
        assert_same(plain(a["summary"]), plain(b["summary"]), "mem.summary")
        assert_same(plain(a["coverage"]), plain(b["coverage"]), "mem.coverage")
        i = ("peak_allocated", "peak_reserved").index(a["metric"])
        assert b["series"]["median"] == fm["_series"][2 * i].tolist()
        assert b["series"]["worst"] == fm["_series"][2 * i + 1].tolist()


@pytest.mark.parametrize("g", STEP, ids=[g["case"] for g in STEP])
def test_fast_oracle_equals_row_oracle_on_goldens(g):
    compare(step_replay_for(g), g["window"])


@pytest.mark.parametrize("scenario,R,S,W,seed", [
    ("balanced", 2, 700, 10_000, 101), ("balanced", 4, 900, 512, 102), ("input_straggler", 8, 400, 10_000, 103),
    ("ragged", 3, 500, 128, 104), ("trend_worsening", 2, 1300, 10_000, 105), ("mem_creep_confirmed", 4, 600, 300, 106),
    ("wait_heavy", 1, 500, 10_000, 107), ("straggler", 6, 300, 10_000, 108), ("balanced", 2, 12_500, 11_000, 109),
])
def test_fast_oracle_equals_row_oracle_seeded(scenario, R, S, W, seed):
    compare(replay.make_step_replay(scenario, R, S, seed=seed), W)


def test_sequential_sum_is_not_pairwise():
    """The property the whole pin rests on: np.add.accumulate is one IEEE add after another."""
    rng = np.random.default_rng(5)
    x = rng.uniform(1.0, 40.0, 200_000)
    acc = 0.0
    for v in x.tolist():
        acc += v
    assert fast_oracle.seq_sum(x) == acc
    assert float(np.sum(x)) != acc  # pairwise: a different rounding history
