"""Pins the CPU oracle's Santorini (no gods) rules and its game-generic MCTS to vectors produced by the reference itself
(tests/golden/santorini_*.npz, made by oracle/gen_golden_santorini.py from a copy of the reference with NB_GODS = 1). Bit-exact."""
import numpy as np

from conftest import MCTS_CONFIGS
from oracle import oracle as O
from oracle.hashnet import hashnet_eval


def test_valid_moves_bit_exact(san_kat):
    k = san_kat
    for cb, v in zip(k['canonical'], k['valids']):
        assert (O.sant_valid_moves(cb, 0) == v).all()
    for b, p, v in zip(k['board'][::3], k['player'][::3], k['valids'][::3]):       # absolute boards, actual mover
        assert (O.sant_valid_moves(b, int(p)) == v).all()


def test_next_state_ended_round_score(san_kat):
    k = san_kat
    for i in range(len(k['action'])):
        nb, npl = O.sant_next_state(k['board'][i], k['player'][i], k['action'][i])
        assert npl == k['next_player'][i]
        assert (nb == k['next_board'][i]).all(), f'ply {i}'
        assert (O.sant_game_ended(nb, npl) == k['ended'][i]).all(), f'ply {i}'
        assert (O.sant_game_ended(k['next_canonical'][i], 0) == k['ended_canonical0'][i]).all()
        assert O.sant_get_round(nb) == k['round'][i]
        assert [O.sant_get_score(nb, 0), O.sant_get_score(nb, 1)] == list(k['score'][i])
    assert (np.abs(k['ended']).sum(axis=1) > 0).sum() == len(np.unique(k['game']))       # every golden game reaches a result


def test_canonical_form(san_kat):
    k = san_kat
    for i in range(len(k['action'])):
        assert (O.sant_canonical(k['board'][i], k['player'][i]) == k['canonical'][i]).all()
        assert (O.sant_canonical(k['next_board'][i], k['next_player'][i]) == k['next_canonical'][i]).all()


def test_symmetries(san_kat):
    k = san_kat
    for i in range(len(k['sym_pi'])):
        s = O.sant_symmetries(k['sym_board'][i], k['sym_pi'][i], k['sym_valids'][i])
        assert len(s) == 8
        for j, (b, p, v) in enumerate(s):
            assert (b == k['sym_out_boards'][i][j]).all(), (i, j)
            assert (p == k['sym_out_pi'][i][j]).all(), (i, j)
            assert (v == k['sym_out_valids'][i][j]).all(), (i, j)


def test_init_game_invariants(san_kat):
    for seed in range(30):
        b = O.sant_init_game(seed)
        assert sorted(b[:, :, 0].ravel().tolist()) == [-2, -1] + [0] * 21 + [1, 2]
        assert not b[:, :, 1].any() and b[:, :, 2].ravel().tolist() == [64, 64] + [0] * 23
    ref = san_kat['init_boards'][0]                                                 # the reference's own initial board has the same form
    assert sorted(ref[:, :, 0].ravel().tolist()) == [-2, -1] + [0] * 21 + [1, 2] and ref[:, :, 2].ravel().tolist() == [64, 64] + [0] * 23


def test_hashnet_162_actions(san_kat):
    for cb, v in zip(san_kat['canonical'][::11], san_kat['valids'][::11]):
        pi, val = O.hashnet(cb, v)
        pi2, val2 = hashnet_eval(cb, v)
        assert (pi == pi2).all() and (val == val2).all()


def _cfg(name, n_sims):
    c = MCTS_CONFIGS[name]
    return O.make_cfg(numMCTSSims=int(n_sims), universes=c['universes'], forced_playouts=c['forced_playouts'], cpuct=c['cpuct'], fpu=c['fpu'],
                      dirichletAlpha=c['dirichletAlpha'], temperature2=c['temperature'][2], net_kind=0, game=O.GAME_SANTORINI), c['noise']


def test_mcts_counts_exact(san_mcts_cases):
    for case in san_mcts_cases:
        cfg, noise = _cfg(str(case['cfg']), case['n_sims'])
        m = O.MCTS(cfg, dirichlet_noise=noise)
        probs, q, full, raw = m.getActionProb(case['root'], temp=1, force_full_search=True, noise=case['noise'])
        assert (raw == case['raw_counts']).all(), str(case['cfg'])
        np.testing.assert_allclose(probs, case['probs'], rtol=0, atol=1e-12)
        assert (q == case['q']).all()
        assert list(m.stats()[:3]) == list(case['summary'])


def test_episode_tree_reuse_exact(san_episode):
    ep = san_episode
    cfg, _ = _cfg('default', ep['n_sims'])
    m = O.MCTS(cfg, dirichlet_noise=False)
    for i in range(len(ep['roots'])):
        probs, q, full, raw = m.getActionProb(ep['roots'][i], temp=1, force_full_search=True)
        assert (raw == ep['raw_counts'][i]).all(), f'ply {i}'
        assert (q == ep['q'][i]).all(), f'ply {i}'
        assert list(m.stats()[:3]) == list(ep['summaries'][i]), f'ply {i}'
