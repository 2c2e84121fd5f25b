"""GPU parity of the Abalone (Belgian daisy) plugin through the C ABI: batched step kernels and the search engine at 3402 actions
against (a) golden vectors produced by the unmodified reference (oracle/gen_golden_abalone.py) and (b) the CPU oracle.
Bit-exact boards / masks / end vectors / symmetries; identical root visit counts (=> policies equal, bar 1e-5); bit-equal q."""
import numpy as np
import pytest

import azg_b200
from azg_b200.mcts import Engine, MCTS
from azg_b200.nnet import HashNetWrapper
from conftest import MCTS_CONFIGS
from oracle import oracle as O
from oracle.hashnet import hashnet_eval

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def game():
    return azg_b200.AbaloneGame()


@pytest.fixture(scope='module')
def hashnet(game):
    return HashNetWrapper(game)


def test_sizes_and_init(game, aba_kat):
    assert game.getBoardSize() == (9, 9, 4) and game.getActionSize() == 3402
    assert (game.getInitBoard() == aba_kat['init_board']).all()


def test_step_kernels_golden(game, aba_kat):
    k = aba_kat
    assert (game.valid_batch(k['canonical']) == k['valids']).all()
    assert (game.valid_batch(k['board'], k['player']) == k['valids']).all()
    nb, npl = game.next_batch(k['board'], k['player'], k['action'], np.zeros(len(k['action']), np.int64))
    assert (npl == k['next_player']).all() and (nb == k['next_board']).all()
    assert (game.ended_batch(k['next_board'], k['next_player']) == k['ended']).all()
    rounds, scores = game.round_score_batch(k['next_board'])
    assert (rounds == k['round']).all() and (scores == k['score']).all()
    assert (game.canonical_batch(k['board'], k['player']) == k['canonical']).all()
    assert (game.canonical_batch(k['next_board'], k['next_player']) == k['next_canonical']).all()


def test_symmetries_golden(game, aba_kat):
    k = aba_kat
    ob, op, ov, ok = game.symmetries_batch(k['sym_board'], k['sym_pi'], k['sym_valids'])
    assert (ok == 12).all()
    assert (ob == k['sym_out_boards']).all() and (op == k['sym_out_pi']).all() and (ov == k['sym_out_valids']).all()


def test_hashnet_3402_actions(game, hashnet, aba_kat):
    b = aba_kat['canonical'][::97]; va = aba_kat['valids'][::97]
    pi, v = hashnet.predict_batch(b, va)
    for i in range(len(b)):
        p2, v2 = hashnet_eval(b[i], va[i])
        assert (pi[i] == p2).all() and (v[i] == v2).all()


def _args(name, n_sims):
    c = MCTS_CONFIGS[name]
    return dict(numMCTSSims=int(n_sims), cpuct=c['cpuct'], fpu=c['fpu'], universes=c['universes'], dirichletAlpha=c['dirichletAlpha'],
                temperature=c['temperature'], forced_playouts=c['forced_playouts'], prob_fullMCTS=1.0, ratio_fullMCTS=5), c['noise']


def test_search_matches_reference(game, hashnet, aba_mcts_cases):
    for case in aba_mcts_cases:
        args, noise = _args(str(case['cfg']), case['n_sims'])
        m = MCTS(game, hashnet, args, dirichlet_noise=noise, node_cap=2048)
        probs, q, full = m.getActionProb(case['root'], temp=1, force_full_search=True, noise=case['noise'])
        assert (m.last_raw_counts == case['raw_counts']).all(), str(case['cfg'])
        np.testing.assert_allclose(np.array(probs), case['probs'], rtol=0, atol=1e-5)
        assert (np.array(q, np.float32) == case['q']).all()
        st = m.engine.stats()
        assert st['sims'] == case['n_sims'] and st['arena_overflows'] == 0
        m.engine.close()


def test_search_1600_sims_matches_reference(game, hashnet, aba_mcts1600_cases):
    """numMCTSSims = 1600 (BASELINE.json configs[4]): root visit counts of the reference, exact."""
    for case in aba_mcts1600_cases:
        args, noise = _args(str(case['cfg']), case['n_sims'])
        m = MCTS(game, hashnet, args, dirichlet_noise=noise, node_cap=4096)
        probs, q, full = m.getActionProb(case['root'], temp=1, force_full_search=True, noise=case['noise'])
        assert (m.last_raw_counts == case['raw_counts']).all(), str(case['cfg'])
        np.testing.assert_allclose(np.array(probs), case['probs'], rtol=0, atol=1e-5)
        assert (np.array(q, np.float32) == case['q']).all()
        st = m.engine.stats()
        assert st['sims'] == 1600 and st['arena_overflows'] == 0 and st['gc_sweeps'] == 0
        m.engine.close()


def test_tree_reuse_episode_matches_reference(game, hashnet, aba_episode):
    ep = aba_episode
    args, _ = _args('default', ep['n_sims'])
    m = MCTS(game, hashnet, args, dirichlet_noise=False, node_cap=1024)      # 30 plies x 100 sims: the exact (tier-1) GC must run
    for i in range(len(ep['roots'])):
        probs, q, full = m.getActionProb(ep['roots'][i], temp=1, force_full_search=True)
        assert (m.last_raw_counts == ep['raw_counts'][i]).all(), f'ply {i}'
        assert (np.array(q, np.float32) == ep['q'][i]).all(), f'ply {i}'
    st = m.engine.stats()
    assert st['arena_overflows'] == 0 and st['gc_sweeps'] == 0 and st['gc_runs'] > 0
    m.engine.close()


def test_random_playouts_vs_oracle(game):
    rng = np.random.default_rng(13)
    n = 96
    boards = np.repeat(game.getInitBoard()[None], n, axis=0).copy()
    players = np.zeros(n, np.int32); alive = np.ones(n, bool)
    for ply in range(130):
        if not alive.any():
            break
        idx = np.flatnonzero(alive)
        valids = game.valid_batch(boards[idx], players[idx])
        acts = np.array([rng.choice(np.flatnonzero(v)) for v in valids], np.int32)
        nb, npl = game.next_batch(boards[idx], players[idx], acts, np.zeros(len(idx), np.int64))
        ended = game.ended_batch(nb, npl)
        for j in range(0, len(idx), 11):
            g = idx[j]
            assert (O.aba_valid_moves(boards[g], int(players[g])) == valids[j]).all()
            ob, onp = O.aba_next_state(boards[g], int(players[g]), int(acts[j]))
            assert (ob == nb[j]).all() and onp == npl[j]
            assert (O.aba_game_ended(nb[j]) == ended[j]).all()
        boards[idx] = nb; players[idx] = npl
        alive[idx] = ~(ended != 0).any(axis=1)
    assert not alive.any(), 'every game ends by round 127'


def test_batched_search_vs_oracle(game, hashnet, aba_kat):
    roots = aba_kat['canonical'][[0, 61, 200, 433]]
    args, _ = _args('universes8', 96)
    eng = Engine(game, hashnet, args, n_games=len(roots), dirichlet_noise=False, node_cap=512)
    counts, raw, q = eng.search(roots)
    c = MCTS_CONFIGS['universes8']
    cfg = O.make_cfg(numMCTSSims=96, universes=c['universes'], forced_playouts=c['forced_playouts'], cpuct=c['cpuct'], fpu=c['fpu'],
                     dirichletAlpha=c['dirichletAlpha'], temperature2=c['temperature'][2], net_kind=0, game=O.GAME_ABALONE)
    for i, r in enumerate(roots):
        probs, oq, full, oraw = O.MCTS(cfg).getActionProb(r, temp=1, force_full_search=True)
        assert (raw[i] == oraw).all() and (q[i] == oq).all()
        assert (counts[i] > 0).sum() <= (raw[i] > 0).sum()           # policy-target pruning only removes mass
    eng.close()


@pytest.mark.parametrize('tag', ['rand', 'shipped'])
def test_v21_forward_golden(game, v21_golden, tag):
    """AbaloneNNet V21 kernel vs the reference's torch CPU fp32 outputs; tolerance 1e-5 absolute on pi and v."""
    from azg_b200.nnet import AbaloneNNetWrapper
    g = v21_golden[tag]
    net = AbaloneNNetWrapper(game, {'nn_version': 21}, state_dict=g['sd'])
    pi, v = net.predict_batch(g['boards'], g['valids'])
    np.testing.assert_allclose(pi, g['pi'], rtol=0, atol=1e-5)
    np.testing.assert_allclose(v, g['v'], rtol=0, atol=1e-5)
    assert (pi[~g['valids']] == 0).all()
    for n in (1, 3, 5):                                        # partial tiles of the 4-leaf CTA tile
        p2, v2 = net.predict_batch(g['boards'][:n], g['valids'][:n])
        np.testing.assert_allclose(p2, g['pi'][:n], rtol=0, atol=1e-5)
        np.testing.assert_allclose(v2, g['v'][:n], rtol=0, atol=1e-5)


def test_v21_forward_large_batch_mixed_tiles(game, v21_golden):
    """720 leaves = full rounds of 4-leaf CTA tiles plus a tail of 2-leaf tiles (net_v21.cuh v21_plan): every replica of the 48 golden
    boards must match the reference's outputs at 1e-5, whichever tile shape evaluated it."""
    from azg_b200.nnet import AbaloneNNetWrapper
    g = v21_golden['shipped']
    net = AbaloneNNetWrapper(game, {'nn_version': 21}, state_dict=g['sd'])
    reps = 15
    pi, v = net.predict_batch(np.concatenate([g['boards']] * reps), np.concatenate([g['valids']] * reps))
    assert pi.shape[0] == 48 * reps
    np.testing.assert_allclose(pi, np.concatenate([g['pi']] * reps), rtol=0, atol=1e-5)
    np.testing.assert_allclose(v, np.concatenate([g['v']] * reps), rtol=0, atol=1e-5)
    np.testing.assert_allclose(pi.sum(1), 1.0, rtol=0, atol=1e-5)


def test_v21_in_the_search_loop(game, v21_golden, aba_kat):
    from azg_b200.nnet import AbaloneNNetWrapper
    sd = v21_golden['shipped']['sd']
    net = AbaloneNNetWrapper(game, {'nn_version': 21}, state_dict=sd)
    roots = aba_kat['canonical'][[0, 150, 400]]
    args, _ = _args('default', 80)
    eng = Engine(game, net, args, n_games=len(roots), node_cap=512)
    counts, raw, q = eng.search(roots)
    cfg = O.make_cfg(numMCTSSims=80, net_kind=3, game=O.GAME_ABALONE)
    exact = 0
    for i, r in enumerate(roots):
        probs, oq, full, oraw = O.MCTS(cfg, blob=O.v21_blob(sd)).getActionProb(r, temp=1, force_full_search=True)
        assert np.abs(raw[i] / raw[i].sum() - probs).max() < 0.08
        exact += int((raw[i] == oraw).all())
    assert exact >= len(roots) - 1
    eng.close()
