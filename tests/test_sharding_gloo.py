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


def _run_ranks(target, extra=(), world=2, attempts=2):
    """Spawn `world` ranks of `target(rank, world, port, *extra, q)` and return what they put on the queue.  A rendezvous
    that fails for reasons outside the code under test (the free port taken in between, a slow first `import torch` in a
    fresh container) is retried once on a new port; the numeric assertions stay with the caller, outside the retry."""
    last = None
    for _ in range(attempts):
        ctx = mp.get_context('spawn')
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=target, args=(r, world, port) + tuple(extra) + (q,)) for r in range(world)]
        for p in procs:
            p.start()
        ok = True
        try:
            res = [q.get(timeout=240) for _ in procs]
        except Exception as e:              # noqa: BLE001 -- queue.Empty: a rank died or never got through the rendezvous
            last, ok, res = e, False, None
        for p in procs:
            p.join(timeout=120)
            if p.is_alive():
                p.kill()
                p.join()
            ok = ok and p.exitcode == 0
            last = last or (None if p.exitcode == 0 else f'rank exited with {p.exitcode}')
        if ok:
            return res
    raise AssertionError(f'the ranks failed {attempts} times: {last!r}')


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
    res = _run_ranks(_worker, (n_total,))
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


def _grad_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from lvae.training import allreduce_flat_gradients
        g = torch.Generator().manual_seed(100 + rank)
        shapes = [(3, 5), (7,), (2, 1, 3, 3), (1,)]
        params = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]
        for p in params[:-1]:                                   # the last one has no gradient on any rank
            p.grad = torch.rand(p.shape, generator=g)
        before = [None if p.grad is None else p.grad.clone() for p in params]
        ptrs = [None if p.grad is None else p.grad.data_ptr() for p in params]
        allreduce_flat_gradients(params, dist.group.WORLD, world)
        in_place = all(p.grad is None or p.grad.data_ptr() == a for p, a in zip(params, ptrs))
        q.put((rank, before, [None if p.grad is None else p.grad.clone() for p in params], in_place))
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_averages_in_place():
    """The training step's single collective (lvae.training.allreduce_flat_gradients; NCCL on the GPUs) on gloo."""
    res = sorted(_run_ranks(_grad_worker), key=lambda t: t[0])
    (_, b0, a0, ip0), (_, b1, a1, ip1) = res
    assert ip0 and ip1
    for x0, x1, y0, y1 in zip(b0, b1, a0, a1):
        if x0 is None:
            assert y0 is None and y1 is None
            continue
        want = (x0 + x1) / 2
        assert torch.allclose(y0, want, atol=1e-7) and torch.equal(y0, y1)


def _bucket_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from lvae.training import FlatLayout, GradientBuckets
        torch.manual_seed(0)                                   # same weights on every rank
        net = torch.nn.Sequential(torch.nn.Linear(40, 70), torch.nn.GELU(), torch.nn.Linear(70, 30), torch.nn.GELU(),
                                  torch.nn.Linear(30, 5))
        unused = torch.nn.Parameter(torch.ones(13))            # never receives a gradient: its bucket is reduced by finish()
        params = list(net.parameters()) + [unused]
        lay = FlatLayout(params)
        flat = lay.new('cpu')
        for p, v in zip(params, lay.views(flat)):
            p.grad = v
        buckets = GradientBuckets(lay, flat, dist.group.WORLD, world, bucket_bytes=4096)      # several small buckets
        x = torch.randn(16, 40, generator=torch.Generator().manual_seed(50 + rank))
        local = torch.autograd.grad(net(x).square().mean(), list(net.parameters()))
        outs = []
        for _ in range(2):                                     # two steps: the hook state resets
            flat.zero_()
            buckets.start()
            net(x).square().mean().backward()
            buckets.finish()
            outs.append([p.grad.clone() for p in net.parameters()])
        views_ok = all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, lay.views(flat)))
        q.put((rank, [g.clone() for g in local], outs, views_ok, len(buckets.ranges), float(unused.grad.abs().max())))
    finally:
        dist.destroy_process_group()


def test_gradient_buckets_average_flat_views_during_backward():
    """lvae.training.GradientBuckets (the data-parallel collective of GraphedTrainStep: gradients are views of one flat
    buffer, reduced bucket by bucket from post-accumulate hooks while the backward runs) on gloo, world size 2."""
    res = sorted(_run_ranks(_bucket_worker), key=lambda t: t[0])
    (_, l0, o0, v0, nb0, u0), (_, l1, o1, v1, nb1, u1) = res
    assert v0 and v1 and nb0 == nb1 and nb0 >= 3 and u0 == 0.0 and u1 == 0.0
    for step in range(2):
        for a, b, g0, g1 in zip(l0, l1, o0[step], o1[step]):
            assert torch.allclose(g0, (a + b) / 2, atol=1e-7) and torch.equal(g0, g1)
