"""CPU, world_size 2 over gloo: ``final_summary``'s time-out rendezvous
(``traceml_b200/summary.py:_rendezvous``).  The reference's call is a file RPC
that returns ``None`` after ``timeout_sec`` (``sdk/summary_client.py:35``); the
collective replacement must fail open the same way when a rank never arrives."""
import os
import sys
import tempfile
import time

import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, init_file, out_dir, absent_rank):
    sys.path.insert(0, os.path.dirname(HERE))
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    from traceml_b200 import summary

    t0 = time.monotonic()
    res = []
    # call 1: everyone arrives (rank 1 late, inside the time-out)
    if rank == 1:
        time.sleep(0.3)
    res.append(summary._rendezvous(5.0, 0.02))
    # call 2: `absent_rank` never calls -> the others give up after the time-out
    if rank != absent_rank:
        t1 = time.monotonic()
        res.append(summary._rendezvous(0.5, 0.02))
        res.append(time.monotonic() - t1)
    else:
        summary._CALLS += 1  # it skipped the call: keep the call counters aligned
        time.sleep(1.0)
        res += [None, 0.0]
    # call 3: everyone again -> the aborted call left nothing behind
    res.append(summary._rendezvous(5.0, 0.02))
    with open(os.path.join(out_dir, f"r{rank}.txt"), "w") as fh:
        fh.write(repr(res))
    dist.barrier()
    dist.destroy_process_group()
    del t0


def test_rendezvous_go_abort_go():
    world = 2
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_worker, args=(world, os.path.join(td, "init"), td, 1), nprocs=world, join=True)
        got = [eval(open(os.path.join(td, f"r{r}.txt")).read()) for r in range(world)]
    assert got[0][0] is True and got[1][0] is True
    assert got[0][1] is False, "rank 0 must fail open when rank 1 never arrives"
    assert 0.4 <= got[0][2] < 3.0, got[0][2]
    assert got[0][3] is True and got[1][3] is True


def _late_worker(rank, world, init_file, out_dir):
    """A rank that arrives after another one already aborted must abort too (unanimity)."""
    sys.path.insert(0, os.path.dirname(HERE))
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    from traceml_b200 import summary

    if rank == 1:
        time.sleep(1.0)  # rank 0 timed out at 0.3 s and published "abort"
    ok = summary._rendezvous(0.3 if rank == 0 else 5.0, 0.02)
    with open(os.path.join(out_dir, f"r{rank}.txt"), "w") as fh:
        fh.write(repr(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_late_rank_obeys_the_abort():
    world = 2
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_late_worker, args=(world, os.path.join(td, "init"), td), nprocs=world, join=True)
        got = [eval(open(os.path.join(td, f"r{r}.txt")).read()) for r in range(world)]
    assert got == [False, False]
