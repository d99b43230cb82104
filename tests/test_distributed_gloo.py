"""World-size-2 CPU test (gloo) of the N>1 host logic: the shard plan, the id broadcast the bench
uses to bootstrap NCCL, and the decomposition itself -- each rank evaluates only its own targets
against all sources (here with the oracle standing in for the GPU), the shards are allgathered,
and the result must equal the single-rank evaluation bit for bit."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys

        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import oracle_lib as O
        from nbodysim_b200 import ic

        # (1) id broadcast as bench.py does for the ncclUniqueId
        idbuf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            idbuf = torch.arange(128, dtype=torch.uint8)
        dist.broadcast(idbuf, 0)
        assert idbuf.tolist() == list(range(128))

        # (2) shard plan + decomposition
        b = ic.plummer(n, seed=5, dims=2)
        npad, start, count = ic.shard_plan(n, world, rank)
        i0, i1 = min(start, n), min(start + count, n)
        mine = np.zeros((count, 2), dtype=np.float32)
        if i1 > i0:
            mine[: i1 - i0] = O.orc_acc(b, 0.05, dims=2, i0=i0, i1=i1)
        parts = [torch.zeros(count, 2) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(mine))
        full = torch.cat(parts).numpy()[:n]
        want = O.orc_acc(b, 0.05, dims=2)
        ok = np.array_equal(full.view(np.uint32), want.view(np.uint32)) and npad % (2048 * world) == 0
        t = torch.tensor([1.0 if ok else 0.0])
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(bool(t.item() == 1.0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [3000, 5000])
def test_world2_shard_decomposition(n):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) is True
