"""CPU, world_size 2 over gloo: the multi-process reduce orchestration
(collectives, step sharding, partial merging, native rule engines) with the
numpy engine double -- the N > 1 host path of ``traceml_b200.reduce`` /
``sections`` without a GPU."""
import os
import sys
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, scenario, S, W, init_file, out_dir, exchange, dense=False):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    from fake_engine import FakeEngine
    import replay
    from traceml_b200 import sections
    from traceml_b200.reduce import TorchDistComm

    recs = replay.make_step_replay(scenario, world, S, seed=3)
    procs = replay.make_proc_replay("overhang", world, 200, seed=3)
    eng = FakeEngine(recs[rank], procs[rank], dense_ok=dense)
    se = sections.SummaryEngine([eng], TorchDistComm(), exchange=exchange,
                                ram_total=replay.PROC_RAM_TOTAL_BYTES, gpu_count=world)
    se.reducer.device = torch.device("cpu")
    res = se.build(W, W)
    red = res.pop("reduce")
    res["_exchanges"] = getattr(se.comm, "n_vec", None)
    torch.save(res, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("scenario,S,W,exchange", [("straggler", 300, 10_000, "nccl"),
                                                   ("ragged", 260, 128, "nccl"),
                                                   ("straggler", 300, 10_000, "a2a"),
                                                   ("ragged", 261, 77, "a2a"),
                                                   ("straggler", 300, 10_000, "dense"),
                                                   ("straggler", 300, 128, "dense")])
def test_two_rank_reduce_matches_oracle(scenario, S, W, exchange):
    """"dense": the engine double reports dense windows, so the lock-step speculation
    (alignment riding in the first exchange) and the dense select path are exercised."""
    from oracle import process_oracle, step_memory_oracle, step_time_oracle
    from helpers import (assert_struct, oracle_mem_rows, oracle_proc_rows, oracle_time_rows, plain,
                         strip_device)
    import replay

    world = 2
    with tempfile.TemporaryDirectory() as td:
        init_file = os.path.join(td, "init")
        mp.spawn(_worker, args=(world, scenario, S, W, init_file, td, "nccl" if exchange == "dense" else exchange,
                                exchange == "dense"), nprocs=world, join=True)
        got = [torch.load(os.path.join(td, f"r{r}.pt"), weights_only=False) for r in range(world)]
    # exchanges per reduce: bounds(+speculative alignment), [alignment], bands
    if scenario == "straggler":  # lock step; "ragged" aligns time and memory separately
        # W < retained rows: memory candidates outnumber the time window -> its own alignment
        want = 3 if exchange != "dense" else (2 if W >= S else 3)
        assert got[0]["_exchanges"] == want, got[0]["_exchanges"]
    for g_ in got:
        g_.pop("_exchanges")
    # every rank computed the identical summary
    assert_struct(plain(got[0]["step_time"]), plain(got[1]["step_time"]), "ranks agree", rel=0.0)
    recs = replay.make_step_replay(scenario, world, S, seed=3)
    ref = step_time_oracle.step_time_section(oracle_time_rows(recs, W), max_rows=W)
    g = got[0]["step_time"]
    assert_struct(plain({k: g["data"][k] for k in g["data"]}),
                  plain({k: ref["data"][k] for k in g["data"]}), "data")
    assert_struct(plain(g["diagnosis"]), plain(ref["diagnosis"]), "diagnosis")
    mref = step_memory_oracle.step_memory_section(oracle_mem_rows(recs), window_size=W,
                                                  gpu_total_bytes=got[0]["step_memory"]["gpu_total_bytes"])
    gd, rd = strip_device(plain(got[0]["step_memory"]["diagnosis"])), strip_device(plain(mref["diagnosis"]))
    assert_struct(gd["primary"], rd["primary"], "mem.primary")
    assert_struct(plain(got[0]["step_memory"]["per_global_rank"]), plain(mref["per_global_rank"]), "mem.rows")
    procs = replay.make_proc_replay("overhang", world, 200, seed=3)
    pref = process_oracle.process_section(oracle_proc_rows(procs, world), max_rows=W)
    assert_struct(plain(got[0]["process"]["primary"]), plain(pref["diagnosis"]["primary"]), "proc.primary")


def _host_worker(rank, world, init_file, out_dir, fake_other_host):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    import socket

    from traceml_b200.reduce import TorchDistComm

    if fake_other_host and rank == 1:
        socket.gethostname = lambda: "some-other-node"
    comm = TorchDistComm()
    first = comm.one_host(torch.device("cpu"))
    socket.gethostname = lambda: f"changed-after-{rank}"   # cached: no second collective, same answer
    second = comm.one_host(torch.device("cpu"))
    torch.save((first, second), os.path.join(out_dir, f"h{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("fake_other_host", [False, True])
def test_one_host_check_is_collective_once_and_agrees(fake_other_host):
    """CUDA-IPC peer loads (the p2p row exchange) need every rank on one host: the check behind
    the ``auto`` choice, over gloo."""
    world = 2
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_host_worker, args=(world, os.path.join(td, "init"), td, fake_other_host), nprocs=world, join=True)
        got = [torch.load(os.path.join(td, f"h{r}.pt"), weights_only=False) for r in range(world)]
    assert got[0] == got[1] == ((not fake_other_host),) * 2
