"""CPU, world_size 2, gloo: the N>1 host path of bench.py -- slab partition, barrier, max-over-ranks
timing reduction and the whole-job throughput arithmetic (no data-path collective exists)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, batch, q):
    sys.path.insert(0, ROOT)
    import importlib

    slab = importlib.import_module("kblas-gpu_b200.slab")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = slab.slab_range(batch, world, rank)
    dist.barrier()
    ms = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)      # pretend per-rank device time
    cnt = torch.tensor([e - b], dtype=torch.int64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    q.put((rank, b, e, float(ms), int(cnt)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [1 << 20, 1001])
def test_two_rank_slabs_and_timing_reduction(batch):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, b0, e0, ms0, c0), (r1, b1, e1, ms1, c1) = res
    assert (b0, e1) == (0, batch) and e0 == b1            # contiguous, complete cover
    assert ms0 == ms1 == 15.0                              # max over ranks
    assert c0 == c1 == batch                               # whole-job unit count
