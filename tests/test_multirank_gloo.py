"""CPU, world_size 2 over gloo: the N>1 logic of bench.py -- disjoint streams per rank (replicas, no
data-path collective), max-over-ranks timing, whole-job aggregation -- plus a real two-rank run of
the host front-end + oracle proving ranks reconstruct different streams deterministically."""
import hashlib
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(ROOT / "tests"))
    import bench
    import p264decoder_b200 as P
    import _oracle as O

    lanes = 2
    digests = []
    for lane in range(lanes):
        syn = P.Synth(6, 4, n_refs=1, seed=bench.stream_seed(rank, lane), intra_pct=5)
        ring = O.OracleFrames(6, 4, 2)
        m = hashlib.md5()
        for _ in range(3):
            for p in ring.recon(syn.next()):
                m.update(p.tobytes())
        digests.append(m.hexdigest())
    local_ms = 10.0 + 5.0 * rank  # rank 1 is the slow one
    dist.barrier()
    ms, = bench.reduce_max([local_ms], dist)
    value = bench.aggregate_value(world, lanes, ms)
    gathered = [None] * world
    dist.all_gather_object(gathered, digests)
    if rank == 0:
        out.put((ms, value, gathered))
    dist.destroy_process_group()


def test_two_ranks_disjoint_streams_and_max_timing():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ms, value, gathered = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ms == 15.0  # max over ranks, not rank 0's own time
    assert value == pytest.approx(2 * 2 / 0.015)
    flat = [d for rank_d in gathered for d in rank_d]
    assert len(set(flat)) == 4  # four different streams were reconstructed


def test_single_process_reduce_is_identity():
    import bench

    assert bench.reduce_max([3.5, 1.0], None) == [3.5, 1.0]
    assert bench.stream_seed(0, 0) != bench.stream_seed(1, 0) != bench.stream_seed(0, 1)
