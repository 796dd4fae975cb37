#!/usr/bin/env python
"""BASELINE config 3 on this engine: DDP MLP where one rank's input pipeline is slow.

    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 \\
        examples/input_straggler_ddp.py [--steps 460] [--sleep-ms 250]

Rank 0's collate function sleeps for every batch (the scenario of the reference's
``examples/input_straggler_ddp_demo.py:25-27``); the traced loop is the stock drop-in

    traceml.init(mode="auto"); with traceml.trace_step(model): ...; traceml.final_summary()

and the end-of-run summary names rank 0 as the INPUT STRAGGLER.  Every rank's records stay in its
own HBM ring; ``final_summary()`` is one collective reduce.
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import traceml_b200 as traceml  # noqa: E402


class Synthetic(torch.utils.data.Dataset):
    def __init__(self, n, dim, classes):
        g = torch.Generator().manual_seed(7)
        self.x = torch.randn(n, dim, generator=g)
        self.y = torch.randint(0, classes, (n,), generator=g)

    def __len__(self):
        return len(self.x)

    def __getitem__(self, i):
        return self.x[i], self.y[i]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=460)
    ap.add_argument("--sleep-ms", type=float, default=250.0)
    ap.add_argument("--slow-rank", type=int, default=0)
    ap.add_argument("--session-root", default=None, help="write final_summary.{json,txt} here (rank 0)")
    args = ap.parse_args()

    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    traceml.init(mode="auto")
    from traceml_b200.runtime import TraceMLRuntime

    sampler = TraceMLRuntime(interval_sec=0.25, native_process_hz=100.0)  # process telemetry, C++ thread
    sampler.start()
    model = torch.nn.Sequential(torch.nn.Linear(1024, 2048), torch.nn.ReLU(),
                                torch.nn.Linear(2048, 2048), torch.nn.ReLU(),
                                torch.nn.Linear(2048, 10)).cuda()
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
    opt = torch.optim.SGD(model.parameters(), lr=1e-2)

    def collate(batch):
        if rank == args.slow_rank:
            time.sleep(args.sleep_ms / 1e3)  # slow INPUT, before the traced step starts
        return torch.utils.data.default_collate(batch)

    batch = 128
    data = Synthetic(batch * args.steps, 1024, 10)
    loader = torch.utils.data.DataLoader(data, batch_size=batch, collate_fn=collate, pin_memory=True)
    for x, y in loader:
        with traceml.trace_step(model):
            x, y = x.to("cuda", non_blocking=True), y.to("cuda", non_blocking=True)
            loss = torch.nn.functional.cross_entropy(model(x), y)
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    sampler.stop()

    summary = traceml.final_summary(session_root=args.session_root)  # collective; rank 0 gets the dict
    if rank == 0:
        st = summary["step_time"]
        diag = st.get("diagnosis") or st.get("primary_diagnosis") or {}
        print(summary["text"])
        print("[demo] step-time diagnosis:", (diag.get("primary") or diag).get("status"))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
