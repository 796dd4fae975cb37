"""Oracle pin, widened: seeded random cases through the UNMODIFIED reference (when it is
importable -- the build container) and the oracle, bit for bit.  The committed golden vectors
are the portable part of the pin; this test re-derives fresh ones wherever the reference lives."""
import os
import random
import sys

import pytest

REF_SRC = "/root/reference/src"
if os.path.isdir(REF_SRC) and REF_SRC not in sys.path:
    sys.path.insert(0, REF_SRC)
pytest.importorskip("traceml.reporting.sections.step_time", reason="reference not importable here")

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

SCENARIOS = ["balanced", "input_straggler", "compute_straggler", "straggler", "input_bound", "wait_heavy",
             "compute_bound", "ragged", "duplicates", "trend_worsening", "mem_creep_early", "mem_imbalance"]


def _cases(n, seed):
    rng = random.Random(seed)
    out = []
    for i in range(n):
        out.append((rng.choice(SCENARIOS), rng.choice([1, 2, 3, 4, 5, 8]), rng.choice([60, 130, 260, 420]),
                    rng.randrange(1000), rng.choice([37, 100, 256, 10_000])))
    return out


@pytest.mark.parametrize("scenario,ranks,steps,seed,window", _cases(10, 2026))
def test_final_summary_sections_oracle_equals_reference(scenario, ranks, steps, seed, window):
    import make_golden as mg

    # run_step_case asserts oracle == reference (floats ==) for step time and step memory
    g = mg.run_step_case(f"fuzz_{scenario}", scenario, ranks, steps, seed, window)
    assert g["step_time"]["data"]["max_rows"] == window


@pytest.mark.parametrize("scenario,ranks,steps,seed,window", _cases(8, 77))
def test_live_views_oracle_equals_reference(scenario, ranks, steps, seed, window):
    import make_live_golden as mlg

    w = min(window, 400)
    mlg.run_live_case(f"fuzz_{scenario}", scenario, ranks, steps, seed, w)      # asserts inside
    mlg.run_live_mem_case(f"fuzz_{scenario}", scenario, ranks, steps, seed, w, None)
