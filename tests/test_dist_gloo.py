"""N>1 host logic on CPU: world_size-2 gloo run of the game sharding and the end-of-iteration example gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import azg_b200
from azg_b200.dist import gather_examples, owner_of, shard_games


def test_shard_games_partition():
    for n, w in ((131072, 8), (16384, 1), (10, 4), (7, 8), (8192, 4)):
        blocks = [shard_games(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
        for r in range(1, w):
            assert blocks[r][0] == blocks[r - 1][0] + blocks[r - 1][1]
        for g in range(0, n, max(1, n // 97)):
            r = owner_of(g, n, w)
            assert blocks[r][0] <= g < blocks[r][0] + blocks[r][1]
    assert shard_games(131072, 3, 8) == (3 * 16384, 16384)          # game g on rank g // 16384


def _examples(rank, n):
    rng = np.random.default_rng(100 + rank)
    return (rng.integers(-5, 9, size=(n, 56, 7), dtype=np.int8), rng.random((n, 81), dtype=np.float32),
            rng.choice(np.float32([-1, 1, 0.01]), size=(n, 2)), rng.random((n, 81)) < 0.4, rng.random((n, 2), dtype=np.float32))


def _worker(rank, world, port, counts, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        want = [np.concatenate([_examples(r, counts[r])[k] for r in range(world)], axis=0) for k in range(5)]
        same = lambda got: all(g.dtype == w.dtype and g.shape == w.shape and (g == w).all() for g, w in zip(got, want))
        got, st = gather_examples(_examples(rank, counts[rank]), return_stats=True)               # every rank gets everything
        ok = same(got) and st['counts'] == list(counts)
        got0 = gather_examples(_examples(rank, counts[rank]), dst=1)                              # point-to-point gather to rank 1 only
        ok = ok and (same(got0) if rank == 1 else all(len(g) == 0 for g in got0))
        tens = tuple(torch.from_numpy(a.view(np.uint8) if a.dtype == np.bool_ else a) for a in _examples(rank, counts[rank]))
        gott = gather_examples(tens)                                                               # tensors in -> tensors out (the device path on GPUs)
        ok = ok and all(isinstance(g, torch.Tensor) for g in gott) and all((g.numpy() == (w.view(np.uint8) if w.dtype == np.bool_ else w)).all() for g, w in zip(gott, want))
        q.put((rank, ok, [g.shape for g in got]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('counts', [(5, 9), (0, 4), (3, 3)])
def test_gather_examples_world2(counts):
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn'); q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, counts, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert res[0][2][0] == (sum(counts), 56, 7)


def test_gather_without_process_group_is_identity():
    ex = _examples(0, 4)
    out = gather_examples(ex)
    assert all((a == b).all() for a, b in zip(ex, out))
