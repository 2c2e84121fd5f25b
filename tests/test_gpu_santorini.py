"""GPU parity of the Santorini (no gods) plugin through the C ABI: batched step kernels and the search engine against
(a) golden vectors produced by the reference (a copy with NB_GODS = 1, oracle/gen_golden_santorini.py) and (b) the CPU oracle.
Bit-exact boards / masks / end vectors / symmetries; identical root visit counts (=> policies equal, bar 1e-5); bit-equal q."""
import numpy as np
import pytest

import azg_b200
from azg_b200.mcts import Engine, MCTS
from azg_b200.nnet import HashNetWrapper
from conftest import MCTS_CONFIGS
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def game():
    return azg_b200.SantoriniGame()


@pytest.fixture(scope='module')
def hashnet(game):
    return HashNetWrapper(game)


def test_sizes(game):
    assert game.getBoardSize() == (5, 5, 3) and game.getActionSize() == 162 and game.getNumberOfPlayers() == 2


def test_step_kernels_golden(game, san_kat):
    k = san_kat
    assert (game.valid_batch(k['canonical']) == k['valids']).all()
    assert (game.valid_batch(k['board'], k['player']) == k['valids']).all()
    nb, npl = game.next_batch(k['board'], k['player'], k['action'], np.full(len(k['action']), 31416, np.int64))
    assert (npl == k['next_player']).all() and (nb == k['next_board']).all()
    assert (game.ended_batch(k['next_board'], k['next_player']) == k['ended']).all()
    assert (game.ended_batch(k['next_canonical']) == k['ended_canonical0']).all()
    rounds, scores = game.round_score_batch(k['next_board'])
    assert (rounds == k['round']).all() and (scores == k['score']).all()
    assert (game.canonical_batch(k['board'], k['player']) == k['canonical']).all()
    assert (game.canonical_batch(k['next_board'], k['next_player']) == k['next_canonical']).all()


def test_symmetries_golden(game, san_kat):
    k = san_kat
    ob, op, ov, ok = game.symmetries_batch(k['sym_board'], k['sym_pi'], k['sym_valids'])
    assert (ok == 8).all()
    assert (ob == k['sym_out_boards']).all() and (op == k['sym_out_pi']).all() and (ov == k['sym_out_valids']).all()


def test_scalar_facade(game, san_kat):
    k = san_kat; i = 23
    b, p = k['board'][i], int(k['player'][i])
    cb = game.getCanonicalForm(b, p)
    assert cb.shape == (5, 5, 3) and cb.dtype == np.int8 and (cb == k['canonical'][i]).all()
    assert (game.getValidMoves(cb, 0) == k['valids'][i]).all()
    nb, npl = game.getNextState(b, p, int(k['action'][i]))
    assert (nb == k['next_board'][i]).all() and npl == k['next_player'][i]
    assert (game.getGameEnded(nb, npl) == k['ended'][i]).all()
    assert len(game.getSymmetries(cb, np.full(162, 1 / 162, np.float32), k['valids'][i])) == 8


def test_random_playouts_vs_oracle(game):
    rng = np.random.default_rng(11)
    n = 384
    boards = game.init_batch(np.arange(1, n + 1, dtype=np.uint64))
    assert (np.sort(boards[:, :, :, 0].reshape(n, -1), axis=1) == np.array([-2, -1] + [0] * 21 + [1, 2], np.int8)).all()
    assert (boards[:, :, :, 2].reshape(n, -1)[:, :2] == 64).all() and len({b.tobytes() for b in boards}) > 370
    players = np.zeros(n, np.int32); alive = np.ones(n, bool)
    for ply in range(110):
        if not alive.any():
            break
        idx = np.flatnonzero(alive)
        valids = game.valid_batch(boards[idx], players[idx])
        assert valids.any(axis=1).all()                       # a player without moves has already lost
        acts = np.array([rng.choice(np.flatnonzero(v)) for v in valids], np.int32)
        nb, npl = game.next_batch(boards[idx], players[idx], acts, np.zeros(len(idx), np.int64))
        ended = game.ended_batch(nb, npl)
        for j in range(0, len(idx), 7):
            g = idx[j]
            assert (O.sant_valid_moves(boards[g], int(players[g])) == valids[j]).all()
            ob, onp = O.sant_next_state(boards[g], int(players[g]), int(acts[j]))
            assert (ob == nb[j]).all() and onp == npl[j]
            assert (O.sant_game_ended(nb[j], int(npl[j])) == ended[j]).all()
        boards[idx] = nb; players[idx] = npl
        alive[idx] = ~(ended != 0).any(axis=1)
    assert not alive.any()


def _args(name, n_sims):
    c = MCTS_CONFIGS[name]
    return dict(numMCTSSims=int(n_sims), cpuct=c['cpuct'], fpu=c['fpu'], universes=c['universes'], dirichletAlpha=c['dirichletAlpha'],
                temperature=c['temperature'], forced_playouts=c['forced_playouts'], prob_fullMCTS=1.0, ratio_fullMCTS=5), c['noise']


def test_search_matches_reference(game, hashnet, san_mcts_cases):
    for case in san_mcts_cases:
        args, noise = _args(str(case['cfg']), case['n_sims'])
        m = MCTS(game, hashnet, args, dirichlet_noise=noise, node_cap=4096)
        probs, q, full = m.getActionProb(case['root'], temp=1, force_full_search=True, noise=case['noise'])
        assert (m.last_raw_counts == case['raw_counts']).all(), str(case['cfg'])
        np.testing.assert_allclose(np.array(probs), case['probs'], rtol=0, atol=1e-5)
        assert (np.array(q, np.float32) == case['q']).all()
        st = m.engine.stats()
        assert st['sims'] == case['n_sims'] and st['arena_overflows'] == 0
        m.engine.close()


def test_tree_reuse_episode_matches_reference(game, hashnet, san_episode):
    ep = san_episode
    args, _ = _args('default', ep['n_sims'])
    m = MCTS(game, hashnet, args, dirichlet_noise=False, node_cap=1536)      # small arena => the exact (tier-1) GC must run
    for i in range(len(ep['roots'])):
        probs, q, full = m.getActionProb(ep['roots'][i], temp=1, force_full_search=True)
        assert (m.last_raw_counts == ep['raw_counts'][i]).all(), f'ply {i}'
        assert (np.array(q, np.float32) == ep['q'][i]).all(), f'ply {i}'
    st = m.engine.stats()
    assert st['arena_overflows'] == 0 and st['gc_sweeps'] == 0
    m.engine.close()


def test_batched_search_vs_oracle(game, hashnet, san_kat):
    roots = san_kat['canonical'][[0, 9, 30, 77, 140, 260, 333, 401]]
    args, _ = _args('shipped', 120)
    eng = Engine(game, hashnet, args, n_games=len(roots), dirichlet_noise=False, node_cap=1024)
    counts, raw, q = eng.search(roots)
    c = MCTS_CONFIGS['shipped']
    cfg = O.make_cfg(numMCTSSims=120, universes=c['universes'], forced_playouts=c['forced_playouts'], cpuct=c['cpuct'], fpu=c['fpu'],
                     dirichletAlpha=c['dirichletAlpha'], temperature2=c['temperature'][2], net_kind=0, game=O.GAME_SANTORINI)
    for i, r in enumerate(roots):
        probs, oq, full, oraw = O.MCTS(cfg).getActionProb(r, temp=1, force_full_search=True)
        assert (raw[i] == oraw).all() and (q[i] == oq).all()
    eng.close()


def test_selfplay_finishes_games(game, hashnet):
    """executeEpisodes on the device with the hash-net: games finish, examples carry legal policies and +-1 results."""
    from azg_b200.coach import Coach
    args = dict(numMCTSSims=32, cpuct=1.25, fpu=0.0, universes=1, dirichletAlpha=-1.0, prob_fullMCTS=1.0, numEps=16)
    c = Coach(game, hashnet, args, n_games=32, seed=3, node_cap=512)
    b, pi, z, va, q = c.raw_examples(16)
    assert len(b) >= 16 and b.shape[1:] == (5, 5, 3)
    assert np.allclose(pi.sum(axis=1), 1.0, atol=1e-5) and (pi[~va] == 0).all()
    assert set(np.unique(z).tolist()) <= {-1.0, 1.0} and (z.sum(axis=1) == 0).all()
    ex = c.augment(b[:4], pi[:4], z[:4], va[:4], q[:4])
    assert len(ex) == 32 and ex[0][0].shape == (5, 5, 3)
    st = c.engine.stats()
    assert st['episodes_finished'] >= 16 and st['arena_overflows'] == 0


@pytest.mark.parametrize('tag', ['rand', 'shipped'])
def test_v89_forward_golden(game, v89_golden, tag):
    """SantoriniNNet V89 kernel vs the reference's torch CPU fp32 outputs; tolerance 1e-5 absolute on pi and v."""
    from azg_b200.nnet import SantoriniNNetWrapper
    g = v89_golden[tag]
    net = SantoriniNNetWrapper(game, {'nn_version': 89}, state_dict=g['sd'])
    pi, v = net.predict_batch(g['boards'], g['valids'])
    np.testing.assert_allclose(pi, g['pi'], rtol=0, atol=1e-5)
    np.testing.assert_allclose(v, g['v'], rtol=0, atol=1e-5)
    assert (pi[~g['valids']] == 0).all()
    p0, v0 = net.predict(g['boards'][5], g['valids'][5])
    assert p0.shape == (162,) and v0.shape == (2,)
    np.testing.assert_allclose(p0, g['pi'][5], rtol=0, atol=1e-5)


def test_v89_forward_vs_oracle_ragged_batches(game, v89_golden, san_kat):
    from azg_b200.nnet import SantoriniNNetWrapper
    sd = v89_golden['rand']['sd']
    net = SantoriniNNetWrapper(game, {'nn_version': 89}, state_dict=sd)
    blob = O.v89_blob(sd)
    for n in (1, 5, 6, 7, 100):                              # partial tiles of the 6-leaf CTA tile
        b = san_kat['canonical'][:n]; va = san_kat['valids'][:n]
        pi, v = net.predict_batch(b, va)
        opi, ov = O.v89_forward(blob, b, va)
        np.testing.assert_allclose(pi, opi, rtol=0, atol=1e-5)
        np.testing.assert_allclose(v, ov, rtol=0, atol=1e-5)


def test_v89_in_the_search_loop(game, v89_golden, san_kat):
    """Real net in the loop: engine vs CPU oracle; nets agree to ~1e-6 so visit counts agree exactly on most roots."""
    from azg_b200.nnet import SantoriniNNetWrapper
    sd = v89_golden['shipped']['sd']
    net = SantoriniNNetWrapper(game, {'nn_version': 89}, state_dict=sd)
    roots = san_kat['canonical'][[0, 50, 120, 300]]
    args, _ = _args('default', 100)
    eng = Engine(game, net, args, n_games=len(roots), node_cap=1024)
    counts, raw, q = eng.search(roots)
    cfg = O.make_cfg(numMCTSSims=100, net_kind=2, game=O.GAME_SANTORINI)
    exact = 0
    for i, r in enumerate(roots):
        probs, oq, full, oraw = O.MCTS(cfg, blob=O.v89_blob(sd)).getActionProb(r, temp=1, force_full_search=True)
        assert np.abs(raw[i] / raw[i].sum() - probs).max() < 0.06
        exact += int((raw[i] == oraw).all())
    assert exact >= len(roots) - 1
    eng.close()
