"""Workload for ncu captures: one rank, W step records resident, N full reduces.
Usage: python profiles/run_reduce.py [W] [reps] [R_virtual]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import replay  # noqa: E402
from traceml_b200 import sections  # noqa: E402
from traceml_b200.engine import Engine  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
R = int(sys.argv[3]) if len(sys.argv) > 3 else 1
torch.cuda.set_device(0)
engines = []
for r in range(R):
    recs = replay.make_step_replay("balanced", R, W, seed=1, only_ranks=[r])[r]
    e = Engine(device=0, rank=r, world=R, ring_slots=W, proc_slots=65536)
    e.load_steps(recs)
    e.load_procs(replay.make_proc_replay("normal", R, 60000, seed=1, only_ranks=[r])[r])
    engines.append(e)
torch.cuda.synchronize()
summ = sections.SummaryEngine(engines, ram_total=replay.PROC_RAM_TOTAL_BYTES, gpu_count=R)
for _ in range(reps):
    out = summ.build(W, 60000, timings=True)
torch.cuda.synchronize()
print("ok", out["step_time"]["diagnosis"]["primary"]["status"], out["reduce"].timings_ms)

if os.environ.get("TML_PYPROF"):
    import cProfile
    import pstats
    import time

    t0 = time.perf_counter()
    for _ in range(20):
        summ.build(W, 60000)
    torch.cuda.synchronize()
    print("wall ms/build", (time.perf_counter() - t0) / 20 * 1e3)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        summ.build(W, 60000)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(25)
