"""world_size-2 test of the batch-sharding host logic on CPU (gloo): ragged shards, gather order, loss assembly."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from lvae import sharding
        g = torch.Generator().manual_seed(0)
        full = torch.rand(n_total, 2, generator=g)              # stand-in for per-image [kl, mse] of the whole batch
        lmb = torch.rand(n_total, generator=g) * 2000 + 16
        mine = sharding.shard_batch(full)                       # what this rank would compute on its GPU
        lo, hi = sharding.shard_range(n_total, rank, world)
        assert mine.shape[0] == hi - lo
        gathered = sharding.gather_per_image(mine, n_total)
        loss, kl, mse = sharding.rate_distortion_summary(gathered, lmb)
        q.put((rank, torch.equal(gathered, full), float(loss), float(kl), float(mse)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_total', [8, 7, 1])
def test_two_rank_shard_gather_equals_single_rank(n_total):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    full = torch.rand(n_total, 2, generator=g)
    lmb = torch.rand(n_total, generator=g) * 2000 + 16
    want = float((full[:, 0].double() + lmb.double() * full[:, 1].double()).mean())
    for rank, same, loss, kl, mse in res:
        assert same, f'rank {rank}: gathered order differs'
        assert loss == want


def test_shard_range_partitions_exactly():
    from lvae.sharding import shard_range
    for n in (0, 1, 5, 8, 13):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
