"""CPU (gloo, world_size 2) tests of the multi-GPU host logic: row sharding, gathering uneven shards, and the flat
gradient all-reduce.  The data path itself has no collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rabi_b200.dist import allreduce_grads, gather_rows, shard_rows

        n = 11  # uneven on purpose
        sl = shard_rows(n)
        full = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)
        got = gather_rows(full[sl] * 1.0, n)
        ok_gather = torch.equal(got, full)
        losses = gather_rows(full[sl, 0].clone(), n)
        ok_gather = ok_gather and torch.equal(losses, full[:, 0])

        torch.manual_seed(0)
        lin = torch.nn.Linear(4, 2)
        x = torch.full((3, 4), float(rank + 1))
        lin(x).sum().backward()
        g_local = lin.weight.grad.clone()
        allreduce_grads(lin.parameters())
        expect = torch.full_like(g_local, 3.0 * (1 + 2) / 2)  # mean over ranks of 3 * (rank + 1)
        ok_grad = torch.allclose(lin.weight.grad, expect)
        q.put((rank, bool(ok_gather), bool(ok_grad), (sl.start, sl.stop)))
    finally:
        dist.destroy_process_group()


def test_row_sharding_gather_and_grad_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(r[1] and r[2] for r in res), res
    assert res[0][3] == (0, 6) and res[1][3] == (6, 11)


def test_shard_rows_partition_properties():
    from rabi_b200.dist import shard_rows

    for n in (0, 1, 7, 10000, 100000):
        for w in (1, 2, 4, 8):
            sl = [shard_rows(n, r, w) for r in range(w)]
            assert sl[0].start == 0 and sl[-1].stop == n
            assert all(a.stop == b.start for a, b in zip(sl[:-1], sl[1:]))
            sizes = [s.stop - s.start for s in sl]
            assert max(sizes) - min(sizes) <= 1
