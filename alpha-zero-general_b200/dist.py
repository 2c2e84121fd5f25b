"""Multi-GPU plumbing: one process per GPU, games sharded by contiguous blocks, ONE exchange per iteration.

The reference has no distributed code; its own scaling method is independent processes (README.md:175-176) and its
self-play games are independent units (Coach.py:93-98). Here game slot `g` of `n_total` lives on rank `g // (n_total/world)`;
every rank holds identical net weights; every RNG stream is keyed by the GLOBAL slot id (`azg_engine_cfg.first_game` =
`shard_games(...)[0]`), so what a slot plays does not depend on the world size (tests/test_gpu_selfplay.py checks that one engine
with 2n slots and two engines with n slots each produce the same multiset of examples). The only exchange is the gather of
finished-game training examples at iteration end (Coach.py:93-103 collects them from the worker threads' queue):

  gather_examples(arrays, dst=None|r)   variable-length gather of the example arrays (boards, pi, z, valids, q). Tensors stay where
      they are: CUDA tensors go through NCCL straight from / into device memory (no host staging, no padding: every rank
      contributes exactly its own byte ranges, placed at its offset of the pre-sized output), numpy / CPU tensors go through the
      group's CPU backend (gloo in the tests). dst=None gives every rank the concatenation (one broadcast per rank and array into
      its slice of the output); dst=r sends only to rank r (point-to-point send/recv batched into one NCCL group), which is what
      the reference's single training process needs and moves world-1 times less data.
  The examples are gathered UN-augmented; getSymmetries (Coach.py:66-69) runs after the gather on the receiving GPU
  (azg_game_symmetries on device pointers), which divides the bytes on the wire by the symmetry count (10-14 / 8 / 12).
"""
import numpy as np


def shard_games(n_total, rank, world):
    """Contiguous block of games owned by `rank`: returns (first_global_game_id, count). Remainder goes to low ranks."""
    if not (0 <= rank < world):
        raise ValueError('rank out of range')
    base, rem = divmod(int(n_total), int(world))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def owner_of(game, n_total, world):
    """Rank that owns global game id `game` under shard_games."""
    base, rem = divmod(int(n_total), int(world))
    cut = rem * (base + 1)
    return game // (base + 1) if game < cut else rem + (game - cut) // max(base, 1)


def gather_examples(arrays, group=None, device=None, dst=None, return_stats=False):
    """Gather variable-length example arrays (a tuple of arrays / tensors sharing their leading dimension) in rank order.

    arrays: numpy arrays, CPU tensors or CUDA tensors. Numpy in -> numpy out (moved through `device` if given, else the group's CPU
    backend); tensors in -> tensors out on the same device. dst=None: every rank returns the full concatenation; dst=r: rank r
    returns it, the other ranks return empty arrays. return_stats adds a dict (counts per rank, bytes received by this rank)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        out = tuple(a if hasattr(a, 'data_ptr') else np.asarray(a) for a in arrays)
        return (out, dict(counts=[int(len(arrays[0]))], bytes_received=0)) if return_stats else out
    world = dist.get_world_size(group); rank = dist.get_rank(group)
    as_numpy = not hasattr(arrays[0], 'data_ptr')
    dev = torch.device(device) if device is not None else (torch.device('cpu') if as_numpy else arrays[0].device)
    local = []
    for a in arrays:
        if as_numpy:
            a = np.ascontiguousarray(a)
            t = torch.from_numpy(a.view(np.uint8) if a.dtype == np.bool_ else a)
            local.append((t.to(dev), a.dtype))
        else:
            local.append((a.contiguous(), None))
    n_local = int(local[0][0].shape[0])
    # 1) counts (one tiny all_gather)
    mine = torch.tensor([n_local], dtype=torch.int64, device=dev)
    lst = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(lst, mine, group=group)
    counts = [int(x.item()) for x in lst]
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64); total = int(offs[-1])
    receiving = dst is None or rank == dst
    outs = [torch.empty((total if receiving else 0,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev) for t, _ in local]
    nbytes = 0
    if dst is None:
        # 2a) every rank's exact range lands in place: one broadcast per (rank, array) into that rank's slice of the output
        works = []
        for r in range(world):
            if counts[r] == 0:
                continue
            for (t, _), o in zip(local, outs):
                sl = o[offs[r]:offs[r + 1]]
                if r == rank:
                    sl.copy_(t)
                else:
                    nbytes += sl.numel() * sl.element_size()
                works.append(dist.broadcast(sl, src=dist.get_global_rank(group, r) if group is not None else r, group=group, async_op=True))
        for w in works:
            w.wait()
    else:
        # 2b) gather to one rank: point-to-point, all transfers batched into a single group (ncclGroupStart/End under NCCL)
        ops = []
        if rank == dst:
            for (t, _), o in zip(local, outs):
                o[offs[rank]:offs[rank + 1]].copy_(t)
            for r in range(world):
                if r == rank or counts[r] == 0:
                    continue
                for o in outs:
                    sl = o[offs[r]:offs[r + 1]]; nbytes += sl.numel() * sl.element_size()
                    ops.append(dist.P2POp(dist.irecv, sl, dist.get_global_rank(group, r) if group is not None else r, group=group))
        elif n_local > 0:
            for t, _ in local:
                ops.append(dist.P2POp(dist.isend, t, dist.get_global_rank(group, dst) if group is not None else dst, group=group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
    if as_numpy:
        res = tuple((o.cpu().numpy().view(np.bool_) if dt == np.bool_ else o.cpu().numpy()) for o, (_, dt) in zip(outs, local))
    else:
        res = tuple(outs)
    return (res, dict(counts=counts, bytes_received=int(nbytes))) if return_stats else res
