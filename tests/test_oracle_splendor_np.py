"""Oracle vs the reference for Splendor with 3 and 4 players: rules bit-exact, MCTS visit counts exact (value rotation
over more than two seats, q = [Qs, -Qs/(np-1), ...]), SplendorNNet V80 on 71 / 88 tokens at 1e-5, Coach.executeEpisode example for
example. Goldens: oracle/gen_golden_splendor_np.py (a scratch copy of the reference with NUMBER_PLAYERS = n)."""
import numpy as np
import pytest

from conftest import MCTS_CONFIGS, assert_examples_equal, load_selfplay_golden, load_splendor_np
from oracle import oracle as O


@pytest.mark.parametrize('n', [3, 4])
def test_rules_bit_exact(n):
    k, _, _ = load_splendor_np(n)
    assert k['board'].shape[1:] == (32 + 10 * n + n * n, 7)
    for i in range(len(k['action'])):
        assert (O.valid_moves(k['canonical'][i], 0, n=n) == k['valids'][i]).all(), i
        assert (O.valid_moves(k['board'][i], int(k['player'][i]), n=n) == k['valids'][i]).all(), i
        nb, npl = O.next_state(k['board'][i], int(k['player'][i]), int(k['action'][i]), int(k['seed'][i]), n=n)
        assert npl == k['next_player'][i] and (nb == k['next_board'][i]).all(), i
        assert (O.game_ended(nb, n=n) == k['ended'][i]).all()
        assert O.get_round(nb) == k['round'][i] and [O.get_score(nb, p, n=n) for p in range(n)] == list(k['score'][i])
        assert (O.canonical(k['board'][i], int(k['player'][i]), n=n) == k['canonical'][i]).all()
        assert (O.canonical(nb, int(npl), n=n) == k['next_canonical'][i]).all()
    for i in range(len(k['sym_k'])):
        s = O.symmetries(k['sym_board'][i], k['sym_pi'][i], k['sym_valids'][i], n=n)
        assert len(s) == k['sym_k'][i]
        for j, (b, p, v) in enumerate(s):
            assert (b == k['sym_out_boards'][i][j]).all() and (p == k['sym_out_pi'][i][j]).all() and (v == k['sym_out_valids'][i][j]).all()


@pytest.mark.parametrize('n', [3, 4])
def test_mcts_counts_exact(n):
    _, cases, _ = load_splendor_np(n)
    for c in cases:
        cf = MCTS_CONFIGS[str(c['cfg'])]
        cfg = O.make_cfg(num_players=n, numMCTSSims=int(c['n_sims']), universes=cf['universes'], forced_playouts=cf['forced_playouts'], cpuct=cf['cpuct'], fpu=cf['fpu'],
                         dirichletAlpha=cf['dirichletAlpha'], temperature2=cf['temperature'][2], net_kind=0)
        m = O.MCTS(cfg, dirichlet_noise=cf['noise'])
        probs, q, full, raw = m.getActionProb(c['root'], temp=1, force_full_search=True, noise=c['noise'])
        assert (raw == c['raw_counts']).all(), str(c['cfg'])
        np.testing.assert_allclose(probs, c['probs'], rtol=0, atol=1e-12)
        assert (q == c['q']).all() and list(m.stats()[:3]) == list(c['summary'])


@pytest.mark.parametrize('n', [3, 4])
def test_v80_forward_and_episode(n):
    _, _, g = load_splendor_np(n)
    pi, v = O.v80_forward(O.v80_blob(g['sd']), g['boards'], g['valids'], n=n)
    np.testing.assert_allclose(pi, g['pi'], rtol=0, atol=1e-5); np.testing.assert_allclose(v, g['v'], rtol=0, atol=1e-5)
    cfg, games = load_selfplay_golden(f'splendor{n}p'); gd = games[0]
    ocfg = O.make_cfg(num_players=n, numMCTSSims=cfg['numMCTSSims'], ratio_fullMCTS=cfg['ratio_fullMCTS'], universes=cfg['universes'], forced_playouts=cfg['forced_playouts'],
                      net_kind=0, cpuct=cfg['cpuct'], fpu=cfg['fpu'], dirichletAlpha=cfg['dirichletAlpha'], prob_fullMCTS=cfg['prob_fullMCTS'], temperature2=cfg['temperature'][2])
    ex = O.execute_episode_inj(ocfg, gd['init'], gd['u_full'], gd['u_move'], gd['chance_seed'], noise=gd['noise'], temperature=cfg['temperature'][:2],
                               tempThreshold=cfg['tempThreshold'])
    assert (ex['actions'] == gd['action']).all() and (ex['full'] == gd['is_full']).all()
    got = []
    for i in range(len(ex['boards'])):
        for b, p, vv in O.symmetries(ex['boards'][i], ex['pi'][i], ex['valids'][i], n=n):
            got.append((b, p, ex['z'][i], vv, ex['q'][i]))
    assert_examples_equal(got, gd)
