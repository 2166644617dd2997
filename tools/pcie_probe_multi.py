#!/usr/bin/env python3
"""Host<->device copy rates of a whole box, independent of the reconstruction engine: every rank (one per GPU, torchrun)
copies pinned buffers H2D, D2H and both at once, all ranks AT THE SAME TIME; rank 0 prints the per-rank and the aggregate
GB/s.  This is the ceiling of the end-to-end (host-buffer) leg of bench.py at N GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe_multi.py
"""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_h2d, n_d2h = 160_000_000, 800_000_000      # one bench step of 256 lanes: 0.16 GB of compact syntax in, 0.8 GB of pictures out
h_in = torch.empty(n_h2d, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n_h2d, dtype=torch.uint8, device="cuda")
h_out = torch.empty(n_d2h, dtype=torch.uint8).pin_memory()
d_out = torch.empty(n_d2h, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d()
    d2h()


def timed(fn, reps=6):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return dt, float(t.item())


rows = []
for name, fn, nbytes in (("H2D alone", h2d, n_h2d), ("D2H alone", d2h, n_d2h), ("H2D + D2H", both, n_h2d + n_d2h)):
    mine, slowest = timed(fn)
    rows.append((name, nbytes / mine / 1e9, world * nbytes / slowest / 1e9))
if rank == 0:
    print(f"{world} GPU(s), all ranks copying at once, pinned host memory ({os.cpu_count()} host cores)")
    for name, per, agg in rows:
        print(f"  {name:10s}: rank 0 {per:6.1f} GB/s, aggregate over the box {agg:7.1f} GB/s (slowest rank)")
if world > 1:
    dist.destroy_process_group()
