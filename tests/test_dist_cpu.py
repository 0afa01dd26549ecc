"""CPU, world_size 2, gloo: the N>1 host logic (shard bounds, replicated RNG draws, output all-gather)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, global_batch, q):
    from disentangledcolorization_b200 import dist as ddist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = ddist.shard_bounds(global_batch, world, rank)
        np.random.seed(130)
        draws = ddist.sharded_init_draws(global_batch, 48, 8, world, rank)
        full = torch.arange(global_batch * 2 * 3 * 4, dtype=torch.float32).view(global_batch, 2, 3, 4)
        got = ddist.gather_outputs(full[lo:hi].clone(), global_batch)
        q.put((rank, lo, hi, draws, torch.equal(got, full)))
    finally:
        dist.destroy_process_group()


def _run(global_batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, global_batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_two_rank_shard_and_gather_equal_and_ragged():
    for gb in (6, 5):
        res = _run(gb)
        np.random.seed(130)
        single = np.stack([np.random.choice(48, 8, replace=False) for _ in range(gb)]).astype(np.int32)
        assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == gb      # contiguous cover
        assert np.array_equal(np.concatenate([res[0][3], res[1][3]]), single)      # same draws as one process
        assert res[0][4] and res[1][4]                                             # gather == single-process concat


def test_shard_bounds_properties():
    from disentangledcolorization_b200.dist import shard_bounds
    for gb in (1, 7, 64, 512):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(gb, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_native_choice_rows_is_numpys_stream():
    """disco_host_choice_rows == `rows` x np.random.choice(S, K, replace=False) on numpy's global generator
    (models/clusterkit.py:107): same index sets, same generator state afterwards, for full and sliced keeps, ragged
    shapes, K == S and S == 1."""
    from disentangledcolorization_b200 import _lib
    for seed in (130, 1, 7, 2 ** 31 - 1):
        for S, K, rows in ((256, 8, 64), (1024, 16, 5), (24, 5, 9), (8, 8, 3), (1, 1, 2), (300, 7, 4)):
            np.random.seed(seed)
            np.random.random(5)                                  # start somewhere inside the 624-word block
            st = np.random.get_state()
            want = np.stack([np.random.choice(S, K, replace=False) for _ in range(rows)])
            nxt = np.random.randint(1 << 30)
            np.random.set_state(st)
            got = _lib.choice_rows(S, K, rows)
            assert got.dtype == np.int32 and np.array_equal(got, want)
            assert np.random.randint(1 << 30) == nxt
            np.random.set_state(st)
            lo, hi = 1, max(1, rows - 1)
            part = _lib.choice_rows(S, K, rows, keep=(lo, hi))
            assert np.array_equal(part, want[lo:hi]) and np.random.randint(1 << 30) == nxt
    import pytest
    with pytest.raises(_lib.DiscoError):
        _lib.choice_rows(4, 5, 1)                                # more clusters than tokens (np.random.choice raises too)
