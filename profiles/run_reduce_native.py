"""Workload for ncu captures of the NATIVE driver (tml_reduce_run): one rank, W step records
resident, N full reduces.  Usage: python profiles/run_reduce_native.py [W] [reps]"""
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
torch.cuda.set_device(0)
recs = replay.make_step_replay("balanced", 1, W, seed=1)[0]
e = Engine(device=0, rank=0, world=1, ring_slots=W, proc_slots=65536)
e.load_steps(recs)
e.load_procs(replay.make_proc_replay("normal", 1, 60000, seed=1)[0])
torch.cuda.synchronize()
summ = sections.SummaryEngine([e], ram_total=replay.PROC_RAM_TOTAL_BYTES, gpu_count=1)
for _ in range(reps):
    out = summ.build(W, 60000)
torch.cuda.synchronize()
red = out["reduce"]
print("ok", out["step_time"]["diagnosis"]["primary"]["status"], "fused_rows", getattr(red, "fused_rows", None), red.timings_ms)
