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
# 12 rounds x 84 steps (~1000 steps): every round first parks ~0.1 s of GPU work on the stream, then
# issues 84 steps' worth of stamps and commits behind it (924 launches: inside the driver's launch
# queue, so what is measured is the calls themselves, not queue back-pressure)
rounds, per = 12, 84
host, waits = [], []
step = 0
for _ in range(rounds):
    torch.cuda.synchronize()
    torch.cuda._sleep(int(0.1 * 1.9e9))
    t1 = time.perf_counter()
    for _s in range(per):
        step += 1
        for ph in (1, 1, 2, 3, 4):
            slot = fast.begin(ph)
            fast.end(ph, slot)
        fast.host(5, 1000)
        fast.commit(step, 0, 0.0)
    host.append(time.perf_counter() - t1)
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    waits.append(time.perf_counter() - t2)
recs, dropped = eng.drain()
print(json.dumps({"steps": step, "stamp_pairs": step * 5, "gpu_busy_in_front_s_per_round": 0.1, "rounds": rounds,
                  "host_issue_s_per_round_max": max(host), "us_per_step_host": sum(host) / step * 1e6,
                  "synchronize_wait_s_per_round_min": min(waits),
                  "records_committed": int(len(recs)) + int(dropped),
                  "host_never_waited": bool(max(host) < 0.05 and min(waits) > 0.04)}))
