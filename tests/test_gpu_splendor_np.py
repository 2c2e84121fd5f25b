"""GPU parity of Splendor with 3 and 4 players through the C ABI: rules and symmetries bit-exact against the reference's
vectors, MCTS visit counts identical (value rotation over more than two seats), SplendorNNet V80 on 71 / 88 tokens (generic token-mixer
kernel) within 1e-5 of the reference's torch outputs, Coach.executeEpisode example for example, and the net in the search loop."""
import numpy as np
import pytest

import azg_b200
from azg_b200.mcts import Engine, MCTS
from azg_b200.nnet import HashNetWrapper, NNetWrapper
from conftest import MCTS_CONFIGS, assert_examples_equal, load_selfplay_golden, load_splendor_np
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('n', [3, 4])
def test_rules_and_symmetries(n):
    game = azg_b200.SplendorGame(n); k, _, _ = load_splendor_np(n)
    assert game.getBoardSize() == (32 + 10 * n + n * n, 7) and game.getActionSize() == 81 and game.getNumberOfPlayers() == n
    assert (game.valid_batch(k['canonical']) == k['valids']).all() and (game.valid_batch(k['board'], k['player']) == k['valids']).all()
    nb, npl = game.next_batch(k['board'], k['player'], k['action'], k['seed'])
    assert (npl == k['next_player']).all() and (nb == k['next_board']).all()
    assert (game.ended_batch(k['next_board'], k['next_player']) == k['ended']).all()
    rounds, scores = game.round_score_batch(k['next_board'])
    assert (rounds == k['round']).all() and (scores == k['score']).all()
    assert (game.canonical_batch(k['board'], k['player']) == k['canonical']).all()
    assert (game.canonical_batch(k['next_board'], k['next_player']) == k['next_canonical']).all()
    ob, op, ov, ok = game.symmetries_batch(k['sym_board'], k['sym_pi'], k['sym_valids'])
    assert (ok == k['sym_k']).all() and (ob == k['sym_out_boards']).all() and (op == k['sym_out_pi']).all() and (ov == k['sym_out_valids']).all()
    boards = game.init_batch(np.arange(1, 65, dtype=np.uint64))
    assert (boards[:, 0, :5] == (5 if n == 3 else 7)).all() and (boards[:, 0, 5] == 5).all()       # bank: 5 / 7 gems per colour, 5 gold


@pytest.mark.parametrize('n', [3, 4])
def test_search_matches_reference(n):
    game = azg_b200.SplendorGame(n); net = HashNetWrapper(game); _, cases, _ = load_splendor_np(n)
    for c in cases:
        cf = MCTS_CONFIGS[str(c['cfg'])]
        args = dict(numMCTSSims=int(c['n_sims']), cpuct=cf['cpuct'], fpu=cf['fpu'], universes=cf['universes'], dirichletAlpha=cf['dirichletAlpha'],
                    temperature=cf['temperature'], forced_playouts=cf['forced_playouts'], prob_fullMCTS=1.0, ratio_fullMCTS=5)
        m = MCTS(game, net, args, dirichlet_noise=cf['noise'], node_cap=4096)
        probs, q, full = m.getActionProb(c['root'], temp=1, force_full_search=True, noise=c['noise'])
        assert (m.last_raw_counts == c['raw_counts']).all(), str(c['cfg'])
        np.testing.assert_allclose(np.array(probs), c['probs'], rtol=0, atol=1e-5)
        assert (np.array(q, np.float32) == c['q']).all() and len(q) == n
        assert m.engine.stats()['arena_overflows'] == 0
        m.engine.close()


@pytest.mark.parametrize('n', [3, 4])
def test_v80_forward_and_search_loop(n):
    game = azg_b200.SplendorGame(n); _, _, g = load_splendor_np(n)
    net = NNetWrapper(game, {'nn_version': 80}, state_dict=g['sd'])
    pi, v = net.predict_batch(g['boards'], g['valids'])
    assert np.abs(pi - g['pi']).max() < 1e-5 and np.abs(v - g['v']).max() < 1e-5 and v.shape == (len(g['boards']), n)
    opi, ov = O.v80_forward(O.v80_blob(g['sd']), g['boards'], g['valids'], n=n)
    assert np.abs(pi - opi).max() < 1e-5 and np.abs(v - ov).max() < 1e-5
    for m_ in (1, 7, 9):
        p2, v2 = net.predict_batch(g['boards'][:m_], g['valids'][:m_])
        assert (p2 == pi[:m_]).all() and (v2 == v[:m_]).all()
    rnd = NNetWrapper(game, {'nn_version': 80}, seed=1)                              # random init: shapes of the n-player net
    boards = game.init_batch(np.arange(1, 17, dtype=np.uint64))
    eng = Engine(game, rnd, dict(numMCTSSims=48, universes=2, prob_fullMCTS=1.0), n_games=16, node_cap=512)
    counts, raw, q = eng.search(boards)
    st = eng.stats(); eng.close()
    assert (raw.sum(axis=1) == 47).all() and st['nn_evals'] > 0 and q.shape == (16, n)


@pytest.mark.parametrize('n', [3, 4])
def test_device_episode_matches_reference_examples(n):
    from test_gpu_selfplay import play_injected, split_by_game
    game = azg_b200.SplendorGame(n)
    cfg, games = load_selfplay_golden(f'splendor{n}p'); gd = games[0]
    ex, st = play_injected(game, cfg, [gd['init']], [gd['u_full']], [gd['u_move']], [gd['chance_seed']], [gd['noise']])
    assert st['episodes_finished'] == 1 and st['arena_overflows'] == 0 and st['examples_dropped'] == 0 and st['gc_sweeps'] == 0
    e0 = split_by_game(ex, [gd['root'][gd['is_full']]])[0]
    got = []
    for i in range(len(e0['boards'])):
        for b, p, vv in O.symmetries(e0['boards'][i], e0['pi'][i], e0['valids'][i], n=n):
            got.append((b, p, e0['z'][i], vv, e0['q'][i]))
    assert_examples_equal(got, gd)
