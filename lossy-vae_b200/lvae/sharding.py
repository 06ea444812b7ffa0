"""Batch sharding across ranks for the rate-distortion path: images are independent units,
so a batch is split into contiguous chunks, every rank evaluates its own chunk with replicated weights, and only
the per-image results are exchanged -- there is no collective on the data path.

Works with any initialised `torch.distributed` backend (NCCL on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous [start, stop) of `n` items owned by `rank`; the first n % world ranks get one extra item."""
    assert 0 <= rank < world and n >= 0
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(batch, rank=None, world=None):
    """The slice of a [B, ...] tensor this rank evaluates."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_range(batch.shape[0], rank, world)
    return batch[lo:hi]


def gather_per_image(values, n_total):
    """values: [n_local, k] per-image results of this rank (e.g. kl, mse) -> [n_total, k] in global image order on
    every rank.  Ragged shards are padded to the largest shard for the all_gather."""
    world, rank = dist.get_world_size(), dist.get_rank()
    k = values.shape[1]
    cap = -(-n_total // world)
    pad = torch.zeros(cap, k, dtype=values.dtype, device=values.device)
    pad[:values.shape[0]] = values
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    out = []
    for r, part in enumerate(parts):
        lo, hi = shard_range(n_total, r, world)
        out.append(part[:hi - lo])
    return torch.cat(out, 0)


def rate_distortion_summary(per_image, lmb):
    """Reference loss convention (qarv/model.py:338-358) from gathered per-image [kl_per_dim, mse]: returns
    (loss, mean kl, mean mse) exactly as a single-rank run over the whole batch computes them."""
    kl, mse = per_image[:, 0].double(), per_image[:, 1].double()
    loss = (kl + lmb.double() * mse).mean()
    return loss, kl.mean(), mse.mean()
