"""GPU parity of the Azul (2 players) plugin and the AzulNNet V84 forward through the C ABI against (a) golden vectors produced by the
UNMODIFIED reference (oracle/gen_golden_azul.py) and (b) the CPU oracle. Bit-exact boards / masks / end vectors / symmetries
(all 120 factory orders); identical root visit counts (=> policies equal, bar 1e-5); bit-equal q; net outputs within 1e-5.
Azul is the game where the same player may move again (the holder of the first-player token starts the next round:
next_player = 0 in the canonical frame => no swap, no value roll, azul/AzulLogicNumba.py:154-158, MCTS.py:176)."""
import numpy as np
import pytest

import azg_b200
from azg_b200.mcts import Engine, MCTS
from azg_b200.nnet import AzulNNetWrapper, HashNetWrapper
from conftest import MCTS_CONFIGS
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def game():
    return azg_b200.AzulGame()


@pytest.fixture(scope='module')
def hashnet(game):
    return HashNetWrapper(game)


def test_sizes(game):
    assert game.getBoardSize() == (23, 6) and game.getActionSize() == 180 and game.getNumberOfPlayers() == 2
    assert game.info.max_symmetries == 120


def test_step_kernels_golden(game, azul_kat):
    k = azul_kat
    assert (game.valid_batch(k['canonical']) == k['valids']).all()
    assert (game.valid_batch(k['board'], k['player']) == k['valids']).all()
    nb, npl = game.next_batch(k['board'], k['player'], k['action'], k['seed'])
    assert (npl == k['next_player']).all() and (nb == k['next_board']).all()
    assert (game.ended_batch(k['next_board'], k['next_player']) == k['ended']).all()
    rounds, scores = game.round_score_batch(k['next_board'])
    assert (rounds == k['round']).all() and (scores == k['score']).all()
    assert (game.canonical_batch(k['board'], k['player']) == k['canonical']).all()
    assert (game.canonical_batch(k['next_board'], k['next_player']) == k['next_canonical']).all()
    # the fixture contains moves after which the SAME player moves again (token holder starts the next round)
    assert ((k['next_player'] == k['player']).sum() > 5)


def test_symmetries_golden(game, azul_kat):
    k = azul_kat
    ob, op, ov, ok = game.symmetries_batch(k['sym_board'], k['sym_pi'], k['sym_valids'])
    assert (ok == 120).all()
    assert (ob == k['sym_out_boards']).all() and (op == k['sym_out_pi']).all() and (ov == k['sym_out_valids']).all()


def test_random_playouts_vs_oracle(game):
    rng = np.random.default_rng(12)
    n = 192
    boards = game.init_batch(np.arange(1, n + 1, dtype=np.uint64))
    assert (boards[:, 1, :5].sum(1) == 80).all() and (boards[:, 4:9, :5].sum(2) == 4).all() and (boards[:, 3, 5] == 1).all() and (boards[:, 0, 2] == 1).all()
    assert len({b.tobytes() for b in boards}) > 180
    players = np.zeros(n, np.int32); alive = np.ones(n, bool)
    for ply in range(130):
        if not alive.any():
            break
        idx = np.flatnonzero(alive)
        valids = game.valid_batch(boards[idx], players[idx])
        assert valids.any(axis=1).all()
        acts = np.array([rng.choice(np.flatnonzero(v)) for v in valids], np.int32)
        seeds = rng.integers(1, 2 ** 31 - 1, len(idx)).astype(np.int64)
        nb, npl = game.next_batch(boards[idx], players[idx], acts, seeds)
        ended = game.ended_batch(nb, npl)
        for j in range(0, len(idx), 5):
            g = idx[j]
            assert (O.azul_valid_moves(boards[g], int(players[g])) == valids[j]).all()
            ob, onp = O.azul_next_state(boards[g], int(players[g]), int(acts[j]), int(seeds[j]))
            assert (ob == nb[j]).all() and onp == npl[j]
            assert (O.azul_game_ended(nb[j]) == ended[j]).all()
        boards[idx] = nb; players[idx] = npl
        alive[idx] = ~(ended != 0).any(axis=1)
    assert not alive.any()


def _args(name, n_sims):
    c = MCTS_CONFIGS[name]
    return dict(numMCTSSims=int(n_sims), cpuct=c['cpuct'], fpu=c['fpu'], universes=c['universes'], dirichletAlpha=c['dirichletAlpha'],
                temperature=c['temperature'], forced_playouts=c['forced_playouts'], prob_fullMCTS=1.0, ratio_fullMCTS=5), c['noise']


def test_search_matches_reference(game, hashnet, azul_mcts_cases):
    assert len(azul_mcts_cases) == 15
    for case in azul_mcts_cases:
        args, noise = _args(str(case['cfg']), case['n_sims'])
        m = MCTS(game, hashnet, args, dirichlet_noise=noise, node_cap=4096)
        probs, q, full = m.getActionProb(case['root'], temp=1, force_full_search=True, noise=case['noise'])
        assert (m.last_raw_counts == case['raw_counts']).all(), str(case['cfg'])
        np.testing.assert_allclose(np.array(probs), case['probs'], rtol=0, atol=1e-5)
        assert (np.array(q, np.float32) == case['q']).all()
        st = m.engine.stats()
        assert st['sims'] == case['n_sims'] and st['arena_overflows'] == 0
        m.engine.close()


def test_tree_reuse_episode_matches_reference(game, hashnet, azul_episode):
    ep = azul_episode
    args, _ = _args(str(ep['cfg']) if 'cfg' in ep else 'default', ep['n_sims'])
    m = MCTS(game, hashnet, args, dirichlet_noise=False, node_cap=2048)
    for i in range(len(ep['roots'])):
        probs, q, full = m.getActionProb(ep['roots'][i], temp=1, force_full_search=True)
        assert (m.last_raw_counts == ep['raw_counts'][i]).all(), f'ply {i}'
        assert (np.array(q, np.float32) == ep['q'][i]).all(), f'ply {i}'
    st = m.engine.stats()
    assert st['arena_overflows'] == 0 and st['gc_sweeps'] == 0
    m.engine.close()


def test_batched_search_vs_oracle(game, hashnet, azul_kat):
    roots = azul_kat['canonical'][[0, 7, 33, 90, 150, 260, 333, 401, 555, 700]]
    args, _ = _args('shipped', 120)
    eng = Engine(game, hashnet, args, n_games=len(roots), dirichlet_noise=False, node_cap=1024)
    counts, raw, q = eng.search(roots)
    c = MCTS_CONFIGS['shipped']
    cfg = O.make_cfg(numMCTSSims=120, universes=c['universes'], forced_playouts=c['forced_playouts'], cpuct=c['cpuct'], fpu=c['fpu'],
                     dirichletAlpha=c['dirichletAlpha'], temperature2=c['temperature'][2], net_kind=0, game=O.GAME_AZUL)
    for i, r in enumerate(roots):
        probs, oq, full, oraw = O.MCTS(cfg).getActionProb(r, temp=1, force_full_search=True)
        assert (raw[i] == oraw).all() and (q[i] == oq).all()
    st = eng.stats(); eng.close()
    assert st['arena_overflows'] == 0


@pytest.mark.parametrize('tag', ['rand', 'shipped'])
def test_v84_forward_matches_reference(game, v84_golden, tag):
    g = v84_golden[tag]
    net = AzulNNetWrapper(game, {'nn_version': 84}, state_dict=g['sd'])
    pi, v = net.predict_batch(g['boards'], g['valids'])
    assert np.abs(pi - g['pi']).max() < 1e-5 and np.abs(v - g['v']).max() < 1e-5       # the reference's torch outputs
    opi, ov = O.v84_forward(O.v84_blob(g['sd']), g['boards'], g['valids'])
    assert np.abs(pi - opi).max() < 1e-5 and np.abs(v - ov).max() < 1e-5               # the oracle on the same weights
    assert (pi[~g['valids']] == 0).all() and np.abs(pi.sum(1) - 1).max() < 1e-5
    for n in (1, 7, 8, 9, 61):                                                           # ragged batches around the 8-leaf tile
        p2, v2 = net.predict_batch(g['boards'][:n], g['valids'][:n])
        assert (p2 == pi[:n]).all() and (v2 == v[:n]).all()
    p1, v1 = net.predict(g['boards'][3], g['valids'][3])
    assert (p1 == pi[3]).all() and (v1 == v[3]).all()


def test_v84_in_the_search_loop(game):
    net = AzulNNetWrapper(game, {'nn_version': 84}, seed=0)
    boards = game.init_batch(np.arange(1, 33, dtype=np.uint64))
    eng = Engine(game, net, dict(numMCTSSims=64, universes=2, prob_fullMCTS=1.0), n_games=32, node_cap=512)
    counts, raw, q = eng.search(boards)
    st = eng.stats(); eng.close()
    assert (raw.sum(axis=1) == 63).all() and st['sims'] == 32 * 64 and st['nn_evals'] > 0 and st['arena_overflows'] == 0
