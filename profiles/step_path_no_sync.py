"""Evidence that the step path never waits for the GPU (no nsys in this image, so the proof is by
construction): a ~0.5 s busy kernel is put on the training stream, then 1 000 traced steps' worth of
step-path calls -- 5 device regions + 1 host region + the commit per step, through the same native
glue trace_step uses -- are issued behind it.  If any of them synchronised with the stream, the
loop would take at least the kernel's 0.5 s; it takes a few tens of milliseconds, and only the
explicit synchronize at the end waits.  Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import traceml_b200 as traceml  # noqa: E402
from traceml_b200 import runtime  # noqa: E402
from traceml_b200.utils import timing  # noqa: E402

torch.cuda.set_device(0)
traceml.init(mode="auto")
eng = runtime.get_engine()
timing._ENG or timing._resolve()
fast = timing._FAST
torch.cuda.synchronize()
eng.drain()
x = torch.zeros(1, device="cuda")
t0 = time.perf_counter()
torch.cuda._sleep(int(0.5 * 1.9e9))          # ~0.5 s of GPU time in front of everything below
t_launch = time.perf_counter() - t0
steps = 1000
t1 = time.perf_counter()
for s in range(1, steps + 1):
    for ph in (1, 1, 2, 3, 4):
        slot = fast.begin(ph)
        fast.end(ph, slot)
    fast.host(5, 1000)
    fast.commit(s, 0, 0.0)
host_s = time.perf_counter() - t1
t2 = time.perf_counter()
torch.cuda.synchronize()
wait_s = time.perf_counter() - t2
recs, dropped = eng.drain()
print(json.dumps({"steps": steps, "stamp_pairs": steps * 5, "gpu_busy_in_front_s": 0.5, "host_issue_s": host_s,
                  "us_per_step_host": host_s / steps * 1e6, "final_synchronize_s": wait_s,
                  "records_committed": int(len(recs)) + int(dropped),
                  "host_never_waited": bool(host_s < 0.25 and wait_s > 0.2)}))
