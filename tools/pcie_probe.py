#!/usr/bin/env python
"""pcie_probe.py -- the host<->device floor under bench.py's e2e number, at N ranks.

Raw pinned-memory H2D, D2H and concurrent (full-duplex) copies of one rank's share of a workload, every rank at the same
time (barrier before, max over ranks), i.e. what jb_process_host could reach at best if rendering cost nothing.
torch is plumbing (pinned memory, streams, torch.distributed).

  python tools/pcie_probe.py [--mib 1500]                                   one GPU
  python -m torch.distributed.run --nproc-per-node 8 tools/pcie_probe.py    all GPUs of the box at once
Prints one JSON line on rank 0: per-rank and aggregate GB/s per direction for h2d, d2h, both.
"""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=1500, help="bytes per direction and rank, MiB (bench C2: 1500; C5 shard: 12000)")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.mib * (1 << 20) // 4
    h_in = torch.empty(n, dtype=torch.float32).pin_memory()
    h_out = torch.empty(n, dtype=torch.float32).pin_memory()
    d = torch.empty(n, dtype=torch.float32, device="cuda")
    d2 = torch.empty(n, dtype=torch.float32, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            d.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d2, non_blocking=True)

    def both():
        h2d()
        d2h()

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            fn()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / args.reps * 1e3
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    gb = n * 4 / 1e9
    out = {"ranks": world, "gb_per_direction_per_rank": gb}
    for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both)):
        ms = timed(fn)
        out[name] = {"ms": ms, "gbs_per_direction_per_rank": gb / ms * 1e3, "gbs_per_direction_aggregate": world * gb / ms * 1e3}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
