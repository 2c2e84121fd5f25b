"""Coach.executeEpisode (Coach.py:37-84) on the device, EXAMPLE FOR EXAMPLE against
  (1) the reference's own executeEpisode (tests/golden/*_selfplay.npz: recorded by running the unmodified reference with every random
      input recorded -- playout-cap coin, Dirichlet draw, move-sampling uniform, chance seed, initial board), and
  (2) the CPU oracle's episode on fresh seeded inputs (more slots, more plies),
for Splendor, Santorini and Abalone, through the C ABI (azg_engine_selfplay_inject + azg_engine_selfplay + azg_engine_examples).
Boards, valids, z, q bit-equal; pi bit-equal (float32(count / total))."""
import numpy as np
import pytest

import azg_b200
from azg_b200.coach import Coach
from azg_b200.mcts import Engine
from azg_b200.nnet import HashNetWrapper
from conftest import assert_examples_equal, load_selfplay_golden
from oracle import oracle as O

pytestmark = pytest.mark.gpu

GAMES = {'splendor': (azg_b200.SplendorGame, O.GAME_SPLENDOR), 'santorini': (azg_b200.SantoriniGame, O.GAME_SANTORINI),
         'abalone': (azg_b200.AbaloneGame, O.GAME_ABALONE), 'azul': (azg_b200.AzulGame, O.GAME_AZUL)}


def engine_args(cfg):
    return dict(numMCTSSims=cfg['numMCTSSims'], cpuct=cfg['cpuct'], fpu=cfg['fpu'], universes=cfg['universes'], dirichletAlpha=cfg['dirichletAlpha'],
                temperature=cfg['temperature'], tempThreshold=cfg['tempThreshold'], forced_playouts=cfg['forced_playouts'],
                prob_fullMCTS=cfg['prob_fullMCTS'], ratio_fullMCTS=cfg['ratio_fullMCTS'])


def oracle_cfg(gid, cfg):
    return O.make_cfg(numMCTSSims=cfg['numMCTSSims'], ratio_fullMCTS=cfg['ratio_fullMCTS'], universes=cfg['universes'],
                      forced_playouts=cfg['forced_playouts'], net_kind=0, cpuct=cfg['cpuct'], fpu=cfg['fpu'], dirichletAlpha=cfg['dirichletAlpha'],
                      prob_fullMCTS=cfg['prob_fullMCTS'], temperature2=cfg['temperature'][2], game=gid)


def play_injected(game, cfg, inits, u_full, u_move, seeds, noises):
    """All episodes concurrently, one per slot. Returns the drained un-augmented example arrays and the engine stats."""
    n = len(inits); A = game.info.action_size
    P = max(len(u) for u in u_full)
    uf = np.zeros((n, P)); um = np.zeros((n, P)); cs = np.ones((n, P), np.int64); nz = np.zeros((n, P, A))
    for g in range(n):
        k = len(u_full[g]); uf[g, :k] = u_full[g]; um[g, :k] = u_move[g]
        if seeds[g] is not None and len(seeds[g]):
            cs[g, :k] = seeds[g]
        for p, x in enumerate(noises[g]):
            k2 = min(len(x), A); nz[g, p, :k2] = x[:k2]
    eng = Engine(game, HashNetWrapper(game), engine_args(cfg), n_games=n, dirichlet_noise=(cfg['dirichletAlpha'] != 0), seed=5,
                 node_cap=8 * cfg['numMCTSSims'] + 512)
    eng.selfplay_inject(np.stack([np.asarray(b).reshape(-1) for b in inits]), uf, um, cs, nz)
    eng.selfplay(min_episodes=n, max_moves=P)
    ex = eng.examples(n * game.info.max_game_len)
    st = eng.stats()
    eng.selfplay_inject()                                         # back to the device RNG: the engine stays usable
    eng.close()
    return ex, st


def split_by_game(ex, want_boards):
    """The ring holds finished games back to back in finishing order: cut it into games by the boards each game must have recorded
    (want_boards[g]: the canonical boards of game g's full-search plies; games may share a prefix, e.g. Abalone's fixed start)."""
    b, pi, z, va, q = ex
    out = {}; cur = 0
    while cur < len(b):
        hit = [g for g in range(len(want_boards)) if g not in out and len(want_boards[g]) > 0 and cur + len(want_boards[g]) <= len(b) and
               (b[cur:cur + len(want_boards[g])].reshape(len(want_boards[g]), -1) == np.asarray(want_boards[g]).reshape(len(want_boards[g]), -1)).all()]
        assert hit, f'the examples from {cur} on are no expected game'
        g = hit[0]; k = len(want_boards[g])
        out[g] = dict(boards=b[cur:cur + k], pi=pi[cur:cur + k], z=z[cur:cur + k], valids=va[cur:cur + k], q=q[cur:cur + k])
        cur += k
    return out


@pytest.mark.parametrize('name', ['splendor', 'santorini', 'abalone', 'azul'])
def test_device_episode_matches_reference_examples(name):
    cls, gid = GAMES[name]; game = cls()
    cfg, games = load_selfplay_golden(name)
    ex, st = play_injected(game, cfg, [g['init'] for g in games], [g['u_full'] for g in games], [g['u_move'] for g in games],
                           [g['chance_seed'] for g in games], [g['noise'] for g in games])
    assert st['episodes_finished'] == len(games) and st['arena_overflows'] == 0 and st['examples_dropped'] == 0 and st['gc_sweeps'] == 0
    per_game = split_by_game(ex, [g['root'][g['is_full']] for g in games])
    assert len(per_game) == len(games)
    for gi, gd in enumerate(games):
        assert_examples_equal(O.augment(gid, per_game[gi]), gd)                # oracle symmetries (pinned by test_oracle_*) on device examples
    # and the device's own getSymmetries kernel through the Coach facade (Coach.py:66-69)
    c = Coach.__new__(Coach); c.game = game
    e0 = per_game[0]
    assert_examples_equal(c.augment(e0['boards'], e0['pi'], e0['z'], e0['valids'], e0['q']), games[0])


@pytest.mark.parametrize('name,n_slots,sims', [('splendor', 12, 40), ('santorini', 16, 64), ('abalone', 5, 32), ('azul', 10, 40)])
def test_device_episode_matches_oracle_on_seeded_inputs(name, n_slots, sims):
    cls, gid = GAMES[name]; game = cls()
    rng = np.random.default_rng(20260 + n_slots)
    cfg = dict(numMCTSSims=sims, cpuct=1.1, fpu=0.1, universes=2 if name in ('splendor', 'azul') else 1, dirichletAlpha=0.4, temperature=[1.0, 0.2, 1.05],
               tempThreshold=8, forced_playouts=(name != 'abalone'), prob_fullMCTS=0.5, ratio_fullMCTS=4)
    P = game.info.max_game_len
    inits = game.init_batch(np.arange(1, n_slots + 1, dtype=np.uint64) * 7919)
    u_full = rng.random((n_slots, P)); u_move = rng.random((n_slots, P)); seeds = rng.integers(1, 2 ** 31 - 1, (n_slots, P))
    Lmax = 128
    noises = [[rng.dirichlet(np.full(Lmax, 0.4)) for _ in range(P)] for _ in range(n_slots)]   # the first L entries are used, whatever L is
    # NOTE a Dirichlet prefix is not renormalised by the reference either when it is injected: both sides consume the same numbers
    want = []
    for g in range(n_slots):
        want.append(O.execute_episode_inj(oracle_cfg(gid, cfg), inits[g], u_full[g], u_move[g], seeds[g], noise=noises[g],
                                          temperature=cfg['temperature'][:2], tempThreshold=cfg['tempThreshold']))
    ex, st = play_injected(game, cfg, list(inits), list(u_full), list(u_move), list(seeds), noises)
    assert st['episodes_finished'] == n_slots and st['arena_overflows'] == 0 and st['examples_dropped'] == 0 and st['gc_sweeps'] == 0
    assert len(ex[0]) == sum(len(w['boards']) for w in want)
    per_game = split_by_game(ex, [w['boards'] for w in want])
    for g, w in enumerate(want):
        if len(w['boards']) == 0:
            continue
        d = per_game[g]
        assert (d['boards'].reshape(len(w['boards']), -1) == w['boards'].reshape(len(w['boards']), -1)).all(), f'game {g}: boards'
        assert (d['pi'] == w['pi']).all(), f'game {g}: pi'
        assert (d['z'] == w['z']).all() and (d['valids'] == w['valids']).all() and (d['q'] == w['q']).all(), f'game {g}'


def test_selfplay_root_noise_is_fresh_every_ply_and_slot():
    """On-device Dirichlet noise in self-play: MCTS.py:187-197 draws a fresh vector for every full search. All slots start from the
    SAME injected board (so priors are identical) with the device drawing the noise; the noise recovered from the stored root priors,
    eta = (P' - 0.75 P) / 0.25 (MCTS.py:190-196), must differ between slots at ply 1 and, within a slot, between ply 1 and ply 2
    (round-1 bug: the stream counter never advanced in self-play, so every ply of a slot drew the same gamma variates)."""
    game = azg_b200.SantoriniGame(); n = 6; A = game.info.action_size
    args = dict(numMCTSSims=2, cpuct=1.25, fpu=0.0, universes=1, dirichletAlpha=0.3, prob_fullMCTS=1.0, temperature=[1.0, 1.0, 1.0], tempThreshold=10)
    eng = Engine(game, HashNetWrapper(game), args, n_games=n, dirichlet_noise=True, seed=3, node_cap=256)
    b0 = game.init_batch(np.array([77], np.uint64))[0]
    P = 4
    eng.selfplay_inject(np.stack([b0.reshape(-1)] * n), np.zeros((n, P)), np.full((n, P), 0.5), np.ones((n, P), np.int64), None)

    def eta(boards_canonical):
        o = eng.node(boards_canonical)
        assert (o['found'] == 1).all()
        out = []
        for i in range(n):
            v = o['Vs'][i]; p0, _ = O.hashnet(boards_canonical[i], v)
            out.append(((o['Ps'][i].astype(np.float64) - 0.75 * p0.astype(np.float64)) / 0.25)[v])
        return out

    eng.selfplay(max_moves=1)
    e1 = eta(np.stack([b0] * n))                                  # player 0 moves first: the canonical root is the initial board itself
    for i in range(n):
        assert abs(e1[i].sum() - 1.0) < 1e-3 and (e1[i] > -1e-6).all()          # it is a Dirichlet sample
        for j in range(i + 1, n):
            assert not np.allclose(e1[i], e1[j], atol=1e-4), 'two slots drew the same root noise'
    boards, players, plies, active = eng.selfplay_state()
    assert (plies == 1).all() and (players == 1).all() and active.all()
    roots2 = np.stack([game.getCanonicalForm(boards[i], int(players[i])) for i in range(n)])
    eng.selfplay(max_moves=1)
    e2 = eta(roots2)
    for i in range(n):
        m = min(len(e1[i]), len(e2[i]))
        a = e1[i][:m] / e1[i][:m].sum(); b = e2[i][:m] / e2[i][:m].sum()
        assert not np.allclose(a, b, atol=1e-3), 'ply 2 reused the gamma variates of ply 1'
    eng.close()


def test_nodes_data_view_matches_reference_root_arrays(mcts_cases):
    """MCTS.nodes_data[s] through azg_engine_node: root Ps (float32) and Qsa (float64) bit-equal to the reference's arrays after a
    recorded search (tests/golden/splendor_mcts.npz root_P / root_Qsa), Nsa = raw counts."""
    from azg_b200.mcts import MCTS
    from conftest import MCTS_CONFIGS
    game = azg_b200.SplendorGame(); net = HashNetWrapper(game)
    done = 0
    for c in mcts_cases:
        cfg = MCTS_CONFIGS[str(c['cfg'])]
        if int(c['n_sims']) > 200:
            continue
        args = dict(numMCTSSims=int(c['n_sims']), cpuct=cfg['cpuct'], fpu=cfg['fpu'], universes=cfg['universes'], dirichletAlpha=cfg['dirichletAlpha'],
                    temperature=cfg['temperature'], forced_playouts=cfg['forced_playouts'], prob_fullMCTS=1.0, ratio_fullMCTS=5)
        m = MCTS(game, net, args, dirichlet_noise=cfg['noise'], node_cap=4096)
        m.getActionProb(c['root'], temp=1, force_full_search=True, noise=c['noise'])
        s = game.stringRepresentation(c['root'])
        assert s in m.nodes_data
        Es, Vs, Ps, Ns, Qsa, Nsa, r, Qs = m.nodes_data[s]
        assert (Ps == c['root_P']).all() and (Qsa == c['root_Qsa']).all() and (Nsa == c['raw_counts']).all()
        assert Ns == int(c['raw_counts'].sum()) and Qs == c['q'][0] and not Es.any()
        m.engine.close(); done += 1
        if done >= 6:
            break
    assert done >= 4


def test_games_do_not_depend_on_the_sharding():
    """per-game RNG keyed by the global slot id. One engine with 8 slots and two engines with 4 slots each
    (first_game 0 and 4: what ranks 0 and 1 of a 2-GPU run would create) play the same games: same multiset of examples."""
    game = azg_b200.SantoriniGame()
    args = dict(numMCTSSims=16, cpuct=1.25, fpu=0.0, universes=1, dirichletAlpha=-1.0, prob_fullMCTS=0.5, ratio_fullMCTS=4)

    def run(n, first):
        eng = Engine(game, HashNetWrapper(game), args, n_games=n, dirichlet_noise=True, seed=77, node_cap=1024, first_game=first)
        eng.selfplay(max_moves=90)
        ex = eng.examples(n * game.info.max_game_len); st = eng.stats(); eng.close()
        assert st['examples_dropped'] == 0 and st['arena_overflows'] == 0
        return ex, st

    def rows(ex):
        return [ex[0][i].tobytes() + ex[1][i].tobytes() + ex[2][i].tobytes() + ex[3][i].tobytes() + ex[4][i].tobytes() for i in range(len(ex[0]))]

    whole, st = run(8, 0)
    a, sa = run(4, 0); b, sb = run(4, 4)
    assert st['episodes_finished'] >= 8 and st['episodes_finished'] == sa['episodes_finished'] + sb['episodes_finished']
    assert sorted(rows(whole)) == sorted(rows(a) + rows(b))
    assert sorted(rows(a)) != sorted(rows(b))                    # and the two shards really play different games


def test_ragged_schedule_plays_the_same_games_as_lock_step(monkeypatch):
    """Playout-cap randomisation (MCTS.py:58-59) gives every search its own budget. The ragged scheduler (k_sp_turn: a slot whose
    budget is spent makes its move at the next launch instead of idling until the longest budget of the ply is done) must play
    exactly the games the lock-step scheduler plays: every example of the lock-step run appears in a longer ragged run, bit for bit,
    and the ragged run really is denser (fewer launches per move)."""
    game = azg_b200.SantoriniGame()
    args = dict(numMCTSSims=20, cpuct=1.25, fpu=0.0, universes=1, dirichletAlpha=-1.0, prob_fullMCTS=0.4, ratio_fullMCTS=4, forced_playouts=True)
    n = 8

    def run(ragged, plies):
        monkeypatch.setenv('AZG_RAGGED', '1' if ragged else '0')
        eng = Engine(game, HashNetWrapper(game), args, n_games=n, dirichlet_noise=True, seed=91, node_cap=1024)
        parts = []
        s0 = eng.stats()
        while eng.stats()['moves_played'] < plies * n:
            left = plies - eng.stats()['moves_played'] // n
            eng.selfplay(max_moves=max(left, 1))
            parts.append(eng.examples(n * game.info.max_game_len))
        st = eng.stats(); eng.close()
        assert st['examples_dropped'] == 0 and st['arena_overflows'] == 0
        ex = tuple(np.concatenate([p[i] for p in parts]) for i in range(5))
        rows = [ex[0][i].tobytes() + ex[1][i].tobytes() + ex[2][i].tobytes() + ex[3][i].tobytes() + ex[4][i].tobytes() for i in range(len(ex[0]))]
        return rows, st

    lock, sl = run(False, 60)
    rag, sr = run(True, 150)
    monkeypatch.delenv('AZG_RAGGED')
    assert sl['episodes_finished'] >= 8 and len(lock) > 50
    from collections import Counter
    missing = Counter(lock) - Counter(rag)
    assert not missing, f'{sum(missing.values())} of {len(lock)} lock-step examples are not in the ragged run'
    # lock-step: 20 launches per ply whatever the budgets; ragged: the mean budget (0.4 * 20 + 0.6 * 5 = 11) per ply
    assert sl['sims'] / sl['moves_played'] < 12.5 and sr['sims'] / sr['moves_played'] < 12.5
    # lock-step: 20 lock-step simulations (select + net + backup launches) per ply of the slowest slot; ragged: ~11 (+ the turn kernel)
    rag_steps_per_move = (sr['kernels_launched'] / 4.0) / (sr['moves_played'] / n)
    assert rag_steps_per_move < 15.0, rag_steps_per_move


@pytest.mark.parametrize('ragged', [False, True])
def test_small_arena_is_trimmed_not_overflowed(ragged, monkeypatch):
    """Tree GC under memory pressure (tree.cuh gc_game): an arena barely larger than one search forces the tier-2 sweep and the tier-3
    trim (the reused tree keeps its breadth-first prefix) on almost every move. No expansion may be lost (arena_overflows = 0: every
    search still stores its numMCTSSims new nodes), every recorded policy is a distribution over legal moves, and the games go on."""
    if not ragged:
        monkeypatch.setenv('AZG_RAGGED', '0')
    game = azg_b200.SplendorGame(); net = HashNetWrapper(game)
    sims = 120
    args = dict(numMCTSSims=sims, cpuct=1.25, fpu=0.0, universes=2, dirichletAlpha=-1.0, temperature=[1.0, 0.1, 1.1], tempThreshold=10,
                forced_playouts=False, prob_fullMCTS=0.5 if ragged else 1.0, ratio_fullMCTS=4, no_mem_optim=False)
    eng = Engine(game, net, args, n_games=24, dirichlet_noise=True, seed=5, node_cap=sims + 40)
    eng.selfplay(min_episodes=24)
    st = eng.stats()
    b, pi, z, va, q = eng.examples(24 * game.info.max_game_len)
    eng.close()
    assert st['arena_overflows'] == 0 and st['examples_dropped'] == 0 and st['episodes_finished'] >= 24
    assert st['gc_sweeps'] > 0 and st['gc_trims'] > 0 and st['max_nodes'] <= st['node_cap'] == sims + 40
    assert len(b) > 100 and np.allclose(pi.sum(axis=1), 1.0, atol=1e-5) and (pi[~va] == 0).all()
