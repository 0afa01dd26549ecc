"""Batch sharding across the GPUs of one box + the single output all-gather.

The forward is independent per image, so ranks take contiguous slices of the global batch with the
weights replicated; the only exchange step is one all-gather of `pred_colors` (NCCL over
NVLink/NVSwitch on GPUs, gloo in the CPU tests).  To match a single-process run of the reference
(whose k-means init consumes `np.random.choice` once per image in batch order), every rank draws the
init index sets of ALL images and keeps its own slice (`sharded_init_draws`).

Limitation (documented, not hidden): the reference re-seeds an EMPTY k-means cluster with `torch.randint` drawn from
the one CPU generator of its single process (clusterkit.py:182).  A shard consumes those draws from offset 0 of its own
process's stream, so a sharded run equals the single-process run bit for bit only while no image of an EARLIER shard hit
an empty cluster (none does on the benchmark inputs; `Engine.kmeans_draws()` reports the count, and
`draws_consumed_before()` below gathers it so a caller can detect the case).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(global_batch, world_size, rank):
    """Contiguous [lo, hi) slice of rank `rank`; the first (global_batch % world_size) ranks get one extra."""
    base, extra = divmod(global_batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sharded_init_draws(global_batch, n_tokens, n_clusters, world_size, rank):
    """np.random.choice(S, K, replace=False) for every image of the GLOBAL batch (same stream on every rank);
    returns this rank's rows as int32 (hi-lo, K)."""
    from . import _lib
    return _lib.choice_rows(n_tokens, n_clusters, global_batch, keep=shard_bounds(global_batch, world_size, rank))


def gather_outputs(local, global_batch=None, group=None):
    """All-gather of a per-rank (b_r, ...) tensor into (sum b_r, ...), in rank order.  One collective when the
    shards are equal (the benchmark configuration); padded all-gather otherwise."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if global_batch is None:
        global_batch = local.shape[0] * world
    sizes = [shard_bounds(global_batch, world, r)[1] - shard_bounds(global_batch, world, r)[0] for r in range(world)]
    assert sizes[rank] == local.shape[0], "local shard does not match shard_bounds()"
    local = local.contiguous()
    if len(set(sizes)) == 1:
        out = local.new_empty((global_batch,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    mx = max(sizes)
    padded = local.new_zeros((mx,) + tuple(local.shape[1:]))
    padded[: local.shape[0]] = local
    buf = local.new_empty((world * mx,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * mx: r * mx + sizes[r]] for r in range(world)], 0)


def draws_consumed_before(local_draws, group=None):
    """Number of empty-cluster `torch.randint` draws consumed by the shards of lower rank in this step (0 on a
    single process).  Non-zero means this rank's empty-cluster re-seeds, if any, would have used later numbers of the
    stream in a single-process run (see the module docstring)."""
    if not dist.is_available() or not dist.is_initialized():
        return 0
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    t = torch.zeros(world, dtype=torch.int64)
    t[rank] = int(local_draws)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = t.to(dev)
    dist.all_reduce(t, group=group)
    return int(t[:rank].sum())
