"""Pins the CPU oracle's Abalone (Belgian daisy) rules and its game-generic MCTS at 3402 actions to vectors produced by the
UNMODIFIED reference (tests/golden/abalone_*.npz, made by oracle/gen_golden_abalone.py). Bit-exact."""
import numpy as np

from conftest import MCTS_CONFIGS
from oracle import oracle as O


def test_init_board(aba_kat):
    assert (O.aba_init_game() == aba_kat['init_board']).all()


def test_valid_moves_bit_exact(aba_kat):
    k = aba_kat
    for i in range(0, len(k['action']), 3):
        assert (O.aba_valid_moves(k['canonical'][i], 0) == k['valids'][i]).all(), i
        assert (O.aba_valid_moves(k['board'][i], int(k['player'][i])) == k['valids'][i]).all(), i


def test_next_state_ended_round_score_canonical(aba_kat):
    k = aba_kat
    for i in range(len(k['action'])):
        nb, npl = O.aba_next_state(k['board'][i], k['player'][i], k['action'][i])
        assert npl == k['next_player'][i] and (nb == k['next_board'][i]).all(), f'ply {i} action {k["action"][i]}'
        assert (O.aba_game_ended(nb) == k['ended'][i]).all()
        assert O.aba_get_round(nb) == k['round'][i] and [O.aba_get_score(nb, 0), O.aba_get_score(nb, 1)] == list(k['score'][i])
        assert (O.aba_canonical(k['board'][i], k['player'][i]) == k['canonical'][i]).all()
        assert (O.aba_canonical(nb, npl) == k['next_canonical'][i]).all()
    assert k['score'].max() >= 4 and (k['ended'] == np.float32(0.001)).any()          # pushes off the board and a drawn game are covered


def test_symmetries(aba_kat):
    k = aba_kat
    for i in range(len(k['sym_pi'])):
        s = O.aba_symmetries(k['sym_board'][i], k['sym_pi'][i], k['sym_valids'][i])
        assert len(s) == 12
        for j, (b, p, v) in enumerate(s):
            assert (b == k['sym_out_boards'][i][j]).all(), (i, j)
            assert (p == k['sym_out_pi'][i][j]).all(), (i, j)
            assert (v == k['sym_out_valids'][i][j]).all(), (i, j)


def _cfg(name, n_sims):
    c = MCTS_CONFIGS[name]
    return O.make_cfg(numMCTSSims=int(n_sims), universes=c['universes'], forced_playouts=c['forced_playouts'], cpuct=c['cpuct'], fpu=c['fpu'],
                      dirichletAlpha=c['dirichletAlpha'], temperature2=c['temperature'][2], net_kind=0, game=O.GAME_ABALONE), c['noise']


def test_mcts_counts_exact(aba_mcts_cases):
    for case in aba_mcts_cases:
        cfg, noise = _cfg(str(case['cfg']), case['n_sims'])
        m = O.MCTS(cfg, dirichlet_noise=noise)
        probs, q, full, raw = m.getActionProb(case['root'], temp=1, force_full_search=True, noise=case['noise'])
        assert (raw == case['raw_counts']).all(), str(case['cfg'])
        np.testing.assert_allclose(probs, case['probs'], rtol=0, atol=1e-12)
        assert (q == case['q']).all()
        assert list(m.stats()[:3]) == list(case['summary'])


def test_mcts_1600_sims_exact(aba_mcts1600_cases):
    """The search length of BASELINE.json configs[4]."""
    assert len(aba_mcts1600_cases) == 2
    test_mcts_counts_exact(aba_mcts1600_cases)


def test_episode_tree_reuse_exact(aba_episode):
    ep = aba_episode
    cfg, _ = _cfg('default', ep['n_sims'])
    m = O.MCTS(cfg, dirichlet_noise=False)
    for i in range(len(ep['roots'])):
        probs, q, full, raw = m.getActionProb(ep['roots'][i], temp=1, force_full_search=True)
        assert (raw == ep['raw_counts'][i]).all(), f'ply {i}'
        assert (q == ep['q'][i]).all(), f'ply {i}'
        assert list(m.stats()[:3]) == list(ep['summaries'][i]), f'ply {i}'
