"""GPU parity of the search engine (select / expand / backup / tree reuse / Dirichlet / forced playouts) against
(a) the reference's MCTS.py run with the deterministic hash-net (golden) and (b) the CPU oracle.
Root visit counts must be identical (=> policies equal to 0 <= 1e-5); q must be bit-equal."""
import numpy as np
import pytest

import azg_b200
from azg_b200.mcts import Engine, MCTS
from azg_b200.nnet import HashNetWrapper, NNetWrapper
from conftest import MCTS_CONFIGS
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def game():
    return azg_b200.SplendorGame()


@pytest.fixture(scope='module')
def hashnet(game):
    return HashNetWrapper(game)


def _args(name, n_sims):
    c = MCTS_CONFIGS[name]
    return dict(numMCTSSims=int(n_sims), cpuct=c['cpuct'], fpu=c['fpu'], universes=c['universes'], dirichletAlpha=c['dirichletAlpha'],
                temperature=c['temperature'], forced_playouts=c['forced_playouts'], prob_fullMCTS=1.0, ratio_fullMCTS=5), c['noise']


def test_single_search_matches_reference(game, hashnet, mcts_cases):
    for case in mcts_cases:
        args, noise = _args(str(case['cfg']), case['n_sims'])
        m = MCTS(game, hashnet, args, dirichlet_noise=noise, node_cap=4096)
        probs, q, full = m.getActionProb(case['root'], temp=1, force_full_search=True, noise=case['noise'])
        assert (m.last_raw_counts == case['raw_counts']).all(), str(case['cfg'])
        np.testing.assert_allclose(np.array(probs), case['probs'], rtol=0, atol=1e-5)
        assert (np.array(q, np.float32) == case['q']).all()
        st = m.engine.stats()
        assert st['sims'] == case['n_sims'] and st['arena_overflows'] == 0
        assert st['expansions'] + st['terminal_hits'] >= case['summary'][0] - case['summary'][1]
        m.engine.close()


def test_batched_search_matches_reference(game, hashnet, mcts_cases):
    """All 'default' cases with 200 sims in ONE engine call: independent trees side by side."""
    cases = [c for c in mcts_cases if str(c['cfg']) == 'default' and c['n_sims'] == 200]
    args, noise = _args('default', 200)
    eng = Engine(game, hashnet, args, n_games=len(cases), dirichlet_noise=noise, node_cap=2048)
    counts, raw, q = eng.search(np.stack([c['root'] for c in cases]))
    for i, c in enumerate(cases):
        assert (raw[i] == c['raw_counts']).all()
        assert (q[i] == c['q']).all()
    eng.close()


@pytest.mark.parametrize('tag', ['A', 'B'])
def test_tree_reuse_episode_matches_reference(game, hashnet, episodes, tag):
    """A whole self-play game: the tree persists between moves (MCTS.py:67-68) and is garbage-collected."""
    ep = episodes[tag]
    args, noise = _args(str(ep['cfg']), ep['n_sims'])
    m = MCTS(game, hashnet, args, dirichlet_noise=noise, node_cap=1024)     # small arena => GC must run and stay exact
    for i in range(len(ep['roots'])):
        nz = ep['noise'][i][:ep['noise_len'][i]]
        probs, q, full = m.getActionProb(ep['roots'][i], temp=1, force_full_search=True, noise=nz)
        assert (m.last_raw_counts == ep['raw_counts'][i]).all(), f'{tag} ply {i}'
        np.testing.assert_allclose(np.array(probs), ep['probs'][i], rtol=0, atol=1e-5)
        assert (np.array(q, np.float32) == ep['q'][i]).all(), f'{tag} ply {i}'
    st = m.engine.stats()
    assert st['gc_runs'] > 0 and st['arena_overflows'] == 0
    m.engine.close()


def test_v80_search_vs_oracle(game, v80_golden, kat):
    """Real net in the loop: GPU engine vs CPU oracle (both mirror the reference's arithmetic). The nets differ by
    float rounding (<=1e-5), so allow rare PUCT flips: policies must agree closely on every root, exactly on most."""
    sd = v80_golden['rand']['sd']
    net = NNetWrapper(game, {'nn_version': 80}, state_dict=sd)
    args, _ = _args('default', 160)
    roots = kat['canonical'][[0, 40, 90, 150, 220, 300]]
    eng = Engine(game, net, args, n_games=len(roots), node_cap=2048)
    counts, raw, q = eng.search(roots)
    cfg = O.make_cfg(numMCTSSims=160, net_kind=1)
    exact = 0
    for i, r in enumerate(roots):
        m = O.MCTS(cfg, blob=O.v80_blob(sd))
        probs, oq, full, oraw = m.getActionProb(r, temp=1, force_full_search=True)
        p = raw[i] / raw[i].sum()
        assert np.abs(p - probs).max() < 0.05
        exact += int((raw[i] == oraw).all())
    assert exact >= len(roots) - 2
    eng.close()


def test_device_buffers_and_stats(game, hashnet, kat):
    torch = pytest.importorskip('torch')
    args, _ = _args('default', 64)
    n = 8
    eng = Engine(game, hashnet, args, n_games=n, node_cap=512)
    roots = np.ascontiguousarray(kat['canonical'][:n]).reshape(n, -1)
    c_host, raw_host, q_host = eng.search(roots)
    eng.reset()
    d_roots = torch.from_numpy(roots).cuda()
    out = (torch.zeros(n, 81, dtype=torch.int32, device='cuda'), torch.zeros(n, 81, dtype=torch.int32, device='cuda'),
           torch.zeros(n, 2, dtype=torch.float32, device='cuda'))
    eng.search(d_roots, out=out)
    torch.cuda.synchronize()
    assert (out[1].cpu().numpy() == raw_host).all() and (out[2].cpu().numpy() == q_host).all()
    st = eng.stats()
    assert st['sims'] == 2 * n * 64 and st['kernels_launched'] > 0
    eng.close()


def test_deep_search_mixed_budgets_vs_oracle(game, hashnet, kat):
    """Long searches (3000 simulations, three universes, forced playouts) from opening, mid- and late-game roots, two consecutive
    searches per tree: deep principal variations exercise the multi-chunk path replay, the cached-choice refresh, the
    all-universe links of deterministic moves and the longest-first work order; every other game runs the short budget
    (numMCTSSims / ratio_fullMCTS), so games leave the work order at different simulations."""
    n_sims, ratio = 3000, 5
    args = dict(numMCTSSims=n_sims, cpuct=0.8, fpu=0.0593, universes=3, dirichletAlpha=0.3, temperature=[1.25, 0.8, 1.1],
                forced_playouts=True, prob_fullMCTS=1.0, ratio_fullMCTS=ratio)
    idx = [0, 55, 130, 210, 330, 450, 560, 610]
    roots = np.ascontiguousarray(kat['canonical'][idx])
    full = np.array([i % 2 == 0 for i in range(len(idx))])
    eng = Engine(game, hashnet, args, n_games=len(idx), node_cap=8192)
    cfg = O.make_cfg(numMCTSSims=n_sims, ratio_fullMCTS=ratio, universes=3, forced_playouts=True, net_kind=0, cpuct=0.8, fpu=0.0593,
                     dirichletAlpha=0.3, prob_fullMCTS=0.0)
    oracles = [O.MCTS(cfg) for _ in idx]
    for rep in range(2):                                        # the second search reuses (and extends) the trees
        counts, raw, q = eng.search(roots, full_search=full)
        for i in range(len(idx)):
            probs, oq, ofull, oraw = oracles[i].getActionProb(roots[i], temp=1, force_full_search=bool(full[i]))
            assert ofull == bool(full[i])
            assert (raw[i] == oraw).all(), f'game {i} search {rep}'
            assert (q[i] == oq).all(), f'game {i} search {rep}'
    st = eng.stats()
    assert st['arena_overflows'] == 0 and st['sims'] == 2 * (4 * n_sims + 4 * (n_sims // ratio))
    eng.close()


def test_path_replay_is_exact_on_deep_trees(game, v80_golden, kat, monkeypatch):
    """With the real net the principal variations get deep (tens of levels, beyond one 32-level replay chunk). The path replay,
    the work order and the persistent warps only change HOW the walk is executed: an engine with the replay switched off
    (AZG_TREE_REPLAY=0) must produce bit-identical visit counts, q values and counters over several consecutive searches."""
    sd = v80_golden['rand']['sd']
    net = NNetWrapper(game, {'nn_version': 80}, state_dict=sd)
    args = dict(numMCTSSims=1200, cpuct=1.25, fpu=0.0, universes=3, dirichletAlpha=-1.0, temperature=[1.0, 0.1, 1.1],
                forced_playouts=False, prob_fullMCTS=1.0, ratio_fullMCTS=5)
    roots = np.ascontiguousarray(kat['canonical'][::10][:64])
    res = []
    for replay in ('1', '0'):
        monkeypatch.setenv('AZG_TREE_REPLAY', replay)
        eng = Engine(game, net, args, n_games=len(roots), node_cap=4096, seed=3)
        out = [tuple(a.copy() for a in eng.search(roots)) for _ in range(2)]
        st = eng.stats(); eng.close()
        res.append((out, st))
    monkeypatch.delenv('AZG_TREE_REPLAY')
    (a, sa), (b, sb) = res
    for x, y in zip(a, b):
        assert (x[1] == y[1]).all() and (x[2] == y[2]).all()
    for k in ('sims', 'node_visits', 'expansions', 'terminal_hits', 'sum_legal_visited'):
        assert sa[k] == sb[k], k
    assert sa['node_visits'] / sa['sims'] > 6                    # the trees really are deep
    assert sa['arena_overflows'] == 0


def test_splendor_selfplay_examples(game, hashnet):
    """Coach.executeEpisodes on the device (Coach.py:37-148) for Splendor with chance nodes: games finish, every example is a
    canonical board with the legal mask the rules give for it (oracle), a policy over legal moves only that sums to 1, a game
    result from the reference's value set, and the run is reproducible from its seed; the facade augments with getSymmetries."""
    from azg_b200.coach import Coach
    args = dict(numMCTSSims=48, cpuct=1.25, fpu=0.0, universes=3, dirichletAlpha=-1.0, prob_fullMCTS=1.0, numEps=12)
    runs = []
    for rep in range(2):
        c = Coach(game, hashnet, args, n_games=24, seed=11, node_cap=512)
        b, pi, z, va, q = c.raw_examples(12)
        st = c.engine.stats()
        runs.append((b.copy(), pi.copy(), z.copy(), va.copy(), q.copy()))
        assert st['episodes_finished'] >= 12 and st['arena_overflows'] == 0
        if rep == 0:
            assert len(b) >= 12 * 10 and b.shape[1:] == (56, 7)
            assert np.allclose(pi.sum(axis=1), 1.0, atol=1e-5) and (pi[~va] == 0).all() and (pi >= 0).all()
            for i in range(0, len(b), 7):
                assert (O.valid_moves(b[i], 0) == va[i]).all()
            zs = set(np.round(np.unique(z).astype(np.float64), 4).tolist())
            assert zs <= {-1.0, 1.0, 0.01, -0.01, 0.0}, zs                 # win / loss / shared win (SplendorLogicNumba.py:221-240)
            assert (np.abs(q) <= 1.0 + 1e-6).all()
            ex = c.augment(b[:3], pi[:3], z[:3], va[:3], q[:3])
            assert 3 * 10 <= len(ex) <= 3 * 14 and ex[0][0].shape == (56, 7)   # 1 + 9 card permutations + <= 4 reserve permutations
            for e_b, e_pi, e_z, e_va, e_q in ex[:12]:
                assert abs(float(e_pi.sum()) - 1.0) < 1e-5 and (e_pi[~e_va.astype(bool)] == 0).all()
        c.engine.close()
    def canon(r):                                                          # the ring is appended to with atomics: compare as a multiset
        rows = [r[0][i].tobytes() + r[1][i].tobytes() + r[2][i].tobytes() + r[3][i].tobytes() + r[4][i].tobytes() for i in range(len(r[0]))]
        return sorted(rows)
    assert canon(runs[0]) == canon(runs[1])                                # same seed, same games
