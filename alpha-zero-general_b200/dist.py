"""Multi-GPU plumbing: one process per GPU, games sharded by contiguous blocks, ONE collective per iteration.

The reference has no distributed code; its own scaling method is independent processes (README.md:175-176) and its
self-play games are independent units (Coach.py:93-98). Here game `g` of `n_total` lives on rank `g // (n_total/world)`;
every rank holds identical net weights; per-game RNG streams are keyed by the GLOBAL game id so results do not depend
on the world size. The only exchange is the gather of finished-game training examples at iteration end
(`gather_examples`), NCCL over NVLink on GPUs, gloo in the CPU tests.
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


def gather_examples(arrays, group=None, device=None):
    """All-gather variable-length example arrays (boards, pi, z, valids, q -- any tuple of numpy arrays sharing their
    leading dimension) from every rank; returns the concatenation in rank order on every rank.

    Counts are exchanged first, then each array is padded to the longest rank's length and moved with ONE
    all_gather per array (NCCL when `device` is a CUDA device, otherwise the group's CPU backend)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return tuple(np.asarray(a) for a in arrays)
    world = dist.get_world_size(group)
    dev = torch.device(device) if device is not None else torch.device('cpu')
    n_local = int(len(arrays[0]))
    mine = torch.tensor([n_local], dtype=torch.int64, device=dev)
    lst = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(lst, mine, group=group)
    counts = [int(x.item()) for x in lst]
    n_max = max(counts) if counts else 0
    out = []
    for a in arrays:
        a = np.ascontiguousarray(a)
        as_bool = a.dtype == np.bool_
        if as_bool:
            a = a.view(np.uint8)
        pad = np.zeros((n_max,) + a.shape[1:], a.dtype)
        pad[:n_local] = a
        t = torch.from_numpy(pad).to(dev)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t, group=group)
        cat = np.concatenate([p[:c].cpu().numpy() for p, c in zip(parts, counts)], axis=0) if n_max else pad
        out.append(cat.view(np.bool_) if as_bool else cat)
    return tuple(out)
