"""The reference's OWN Coach.executeEpisode (Coach.py:37-84), unmodified, driven over this package's facades
(`SplendorGame`, `MCTS`, `HashNetWrapper`) -- SURVEY 8b: "the reference's own Coach.learn, Arena, pit.py must be able to run unmodified
against those facades". There is no machine with both the reference and a GPU, so the C ABI underneath the facades is replaced by a
FAKE `lib` that answers every azg_* call from the CPU oracle (test infrastructure; the product path has no such fallback). What is
checked is the facade layer itself: argument / return types and shapes, aliasing rules, the tuple layout the reference's loop expects.
Needs /root/reference (build container only)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='needs the reference tree (build container only)')


def _arr(addr, shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    return np.frombuffer((C.c_uint8 * n).from_address(addr), dtype=dtype).reshape(shape)


class FakeLib:
    """Oracle-backed stand-in for libazg_b200.so (Splendor, 2 players, hash-net): same entry points, same buffers."""
    S, A, NP, K = 392, 81, 2, 14

    def __init__(self):
        from oracle import oracle as O
        self.O = O; self.engines = {}; self.next_h = 1; self.calls = {}

    def _count(self, name): self.calls[name] = self.calls.get(name, 0) + 1
    def azg_abi_version(self): return 4
    def azg_last_error(self): return b''
    def azg_device_count(self): return 1

    def azg_game_info(self, gid, npl, out):
        gi = out._obj
        gi.game_id, gi.num_players, gi.state_rows, gi.state_cols, gi.state_depth = 1, 2, 56, 7, 1
        gi.state_bytes, gi.action_size, gi.max_symmetries, gi.max_game_len = self.S, self.A, self.K, 124
        return 0

    def azg_game_init(self, gid, npl, n, seeds, boards, stream):
        self._count('init'); s = _arr(seeds, (n,), np.uint64); b = _arr(boards, (n, 56, 7), np.int8)
        for i in range(n):
            b[i] = self.O.init_game(int(s[i]) & 0x7FFFFFFF)
        return 0

    def azg_game_valid(self, gid, npl, n, boards, players, mask, stream):
        self._count('valid'); b = _arr(boards, (n, 56, 7), np.int8); m = _arr(mask, (n, self.A), np.uint8)
        pl = _arr(players, (n,), np.int32) if players else np.zeros(n, np.int32)
        for i in range(n):
            m[i] = self.O.valid_moves(b[i], int(pl[i]))
        return 0

    def azg_game_next(self, gid, npl, n, boards, players, actions, seeds, keys, out, out_np, stream):
        self._count('next'); b = _arr(boards, (n, 56, 7), np.int8); o = _arr(out, (n, 56, 7), np.int8); onp = _arr(out_np, (n,), np.int32)
        pl = _arr(players, (n,), np.int32); ac = _arr(actions, (n,), np.int32); sd = _arr(seeds, (n,), np.int64); ks = _arr(keys, (n,), np.uint64)
        for i in range(n):
            nb, nxt = self.O.next_state(b[i], int(pl[i]), int(ac[i]), int(sd[i]), rng_seed=int(ks[i]) & 0x7FFFFFFF)
            o[i] = nb; onp[i] = nxt
        return 0

    def azg_game_ended(self, gid, npl, n, boards, next_players, out, stream):
        self._count('ended'); b = _arr(boards, (n, 56, 7), np.int8); o = _arr(out, (n, 2), np.float32)
        for i in range(n):
            o[i] = self.O.game_ended(b[i])
        return 0

    def azg_game_canonical(self, gid, npl, n, boards, players, out, stream):
        self._count('canonical'); b = _arr(boards, (n, 56, 7), np.int8); o = _arr(out, (n, 56, 7), np.int8); pl = _arr(players, (n,), np.int32)
        for i in range(n):
            o[i] = self.O.canonical(b[i], int(pl[i]))
        return 0

    def azg_game_round_score(self, gid, npl, n, boards, rounds, scores, stream):
        b = _arr(boards, (n, 56, 7), np.int8)
        if rounds:
            _arr(rounds, (n,), np.int32)[:] = [self.O.get_round(x) for x in b]
        if scores:
            _arr(scores, (n, 2), np.int32)[:] = [[self.O.get_score(x, 0), self.O.get_score(x, 1)] for x in b]
        return 0

    def azg_game_symmetries(self, gid, npl, n, boards, pi, mask, ob, opi, om, ok, stream):
        self._count('symmetries'); b = _arr(boards, (n, 56, 7), np.int8); p = _arr(pi, (n, self.A), np.float32); m = _arr(mask, (n, self.A), np.uint8)
        o_b = _arr(ob, (n, self.K, 56, 7), np.int8); o_p = _arr(opi, (n, self.K, self.A), np.float32); o_m = _arr(om, (n, self.K, self.A), np.uint8); o_k = _arr(ok, (n,), np.int32)
        o_b[:] = 0; o_p[:] = 0; o_m[:] = 0
        for i in range(n):
            sy = self.O.symmetries(b[i], p[i], m[i])
            o_k[i] = len(sy)
            for k, (sb, sp, sv) in enumerate(sy):
                o_b[i, k] = sb; o_p[i, k] = sp; o_m[i, k] = sv
        return 0

    def azg_net_create(self, kind, gid, npl, w, nw, out):
        assert kind == 0; out._obj.value = 7; return 0
    def azg_net_destroy(self, h): return 0

    def azg_engine_create(self, cfg, net, out):
        c = cfg._obj; O = self.O
        ocfg = O.make_cfg(numMCTSSims=c.numMCTSSims, ratio_fullMCTS=c.ratio_fullMCTS, universes=c.universes, forced_playouts=bool(c.forced_playouts),
                          net_kind=0, cpuct=c.cpuct, fpu=c.fpu, dirichletAlpha=c.dirichletAlpha, prob_fullMCTS=c.prob_fullMCTS, temperature2=c.temperature[2])
        h = self.next_h; self.next_h += 1
        self.engines[h] = O.MCTS(ocfg, None, dirichlet_noise=bool(c.dirichlet_noise), seed=int(c.seed) + 1)
        out._obj.value = h; return 0

    def azg_engine_destroy(self, h): self.engines.pop(getattr(h, 'value', h), None); return 0
    def azg_engine_reset(self, h, game): self.engines[getattr(h, 'value', h)].reset(); return 0

    def azg_engine_search(self, h, n, roots, full, noise, counts, raw, q, stream):
        self._count('search'); m = self.engines[getattr(h, 'value', h)]
        r = _arr(roots, (n, 56, 7), np.int8); f = _arr(full, (n,), np.uint8) if full else np.ones(n, np.uint8)
        assert n == 1
        probs, oq, is_full, oraw = m.getActionProb(r[0], temp=1, force_full_search=bool(f[0]))
        # counts after policy-target pruning: probs are counts / sum at temp 1; give back integers with the same ratios
        tot = int(oraw.sum()); c = np.rint(np.asarray(probs) * (tot if not m.cfg.forced_playouts else 1)).astype(np.int32) if not m.cfg.forced_playouts else None
        _arr(counts, (n, self.A), np.int32)[0] = oraw if c is None else c
        if raw: _arr(raw, (n, self.A), np.int32)[0] = oraw
        if q: _arr(q, (n, 2), np.float32)[0] = oq
        return 0


@pytest.fixture()
def fake(monkeypatch):
    from azg_b200 import lib                                            # the package's own module object (not a second import under the alias)
    f = FakeLib()
    monkeypatch.setattr(lib, 'load', lambda: f)
    monkeypatch.setattr(lib, '_lib', f, raising=False)
    return f


def test_reference_execute_episode_runs_over_the_facades(fake):
    sys.path[:0] = [os.path.join(os.path.dirname(__file__), '..', 'oracle', 'ref_shim'), REF]
    os.environ.setdefault('NUMBA_CACHE_DIR', '/tmp/numba_cache')
    import Coach as ref_coach                                          # the reference's Coach.py, unmodified
    import azg_b200
    from azg_b200.mcts import MCTS
    from azg_b200.nnet import HashNetWrapper
    from azg_b200.utils import dotdict
    game = azg_b200.SplendorGame()
    game._seed_ctr = 20261017                                           # the facade seeds its real-move RNG keys from OS entropy: fixed here, the test plays one known game
    net = HashNetWrapper(game)
    args = dotdict(numMCTSSims=12, cpuct=1.25, fpu=0.0, universes=1, dirichletAlpha=0.3, temperature=[1.0, 0.1, 1.1], tempThreshold=10,
                   prob_fullMCTS=0.5, ratio_fullMCTS=3, forced_playouts=False, no_mem_optim=False, no_compression=True)
    mcts = MCTS(game, net, args, dirichlet_noise=True, seed=3)
    coach = ref_coach.Coach.__new__(ref_coach.Coach)                   # __init__ builds torch nets; executeEpisode only needs these
    coach.game = game; coach.args = args; coach.mcts = mcts; coach.nb_threads = 1
    np.random.seed(1)
    examples = coach.executeEpisode(mcts, game)                        # <- reference code calling the facades
    assert len(examples) >= 10 and fake.calls['search'] >= 20 and fake.calls['symmetries'] >= 1 and fake.calls['next'] == fake.calls['search']
    for b, pi, z, valids, q in examples[:40]:
        assert isinstance(b, np.ndarray) and b.dtype == np.int8 and b.shape == (56, 7)
        assert pi.dtype == np.float32 and pi.shape == (81,) and abs(float(pi.sum()) - 1.0) < 1e-5
        assert z.dtype == np.float32 and z.shape == (2,) and all(min(abs(abs(float(x)) - 1.0), abs(abs(float(x)) - 0.01)) < 1e-6 for x in z)   # win / loss, or the 0.01 of a tie (a game that ran into the round limit)
        assert valids.shape == (81,) and valids.dtype == np.bool_ and (pi[~valids] == 0).all()
        assert len(q) == 2 and abs(float(q[0]) + float(q[1])) < 1e-6
    # the reference's example pipeline accepts them: Coach.py:172-176 (valid-move statistics) and the on-disk format
    assert 1 < sum(sum(x[3]) for x in examples) / len(examples) < 81
    from azg_b200 import formats as F
    b, pi, z, va, q = F.examples_to_arrays(examples)
    assert b.shape == (len(examples), 56, 7) and q.shape == (len(examples), 2)
