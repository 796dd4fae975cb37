"""Workload for an ncu capture of K4 under REAL peer loads: torchrun, one rank per GPU, every rank
holds W step records, the native driver reduces with the fused NVLink peer-load exchange ("p2p").
Launch through profiles/ncu_rank0.sh so that only rank 0 runs under ncu.
Usage: ... profiles/run_reduce_p2p.py [W] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import replay  # noqa: E402
from traceml_b200 import sections  # noqa: E402
from traceml_b200.engine import Engine  # noqa: E402
from traceml_b200.reduce import TorchDistComm  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
recs = replay.make_step_replay("balanced", world, W, seed=1, only_ranks=[rank])[rank]
e = Engine(device=local, rank=rank, world=world, ring_slots=W, proc_slots=65536)
e.load_steps(recs)
e.load_procs(replay.make_proc_replay("normal", world, 60000, seed=1, only_ranks=[rank])[rank])
torch.cuda.synchronize()
summ = sections.SummaryEngine([e], TorchDistComm(), exchange="p2p", ram_total=replay.PROC_RAM_TOTAL_BYTES,
                              gpu_count=world)
for _ in range(reps):
    out = summ.build(W, 60000)
torch.cuda.synchronize()
if rank == 0:
    print("ok", out["step_time"]["diagnosis"]["primary"]["status"], out["reduce"].exchange, out["reduce"].timings_ms)
dist.barrier()
e.close()
dist.destroy_process_group()
