"""Pins the oracle's MCTS (oracle/azg_oracle.c: search / getActionProb) to the reference's MCTS.py run
with the deterministic hash-net (tests/golden/splendor_mcts.npz, splendor_episode_*.npz)."""
import numpy as np

from conftest import MCTS_CONFIGS
from oracle import oracle as O


def _cfg(name, n_sims):
    c = MCTS_CONFIGS[name]
    return O.make_cfg(numMCTSSims=int(n_sims), universes=c['universes'], forced_playouts=c['forced_playouts'],
                      cpuct=c['cpuct'], fpu=c['fpu'], dirichletAlpha=c['dirichletAlpha'], temperature2=c['temperature'][2],
                      net_kind=0), c['noise']


def test_single_search_counts_exact(mcts_cases):
    for case in mcts_cases:
        cfg, noise = _cfg(str(case['cfg']), case['n_sims'])
        m = O.MCTS(cfg, dirichlet_noise=noise)
        probs, q, full, raw = m.getActionProb(case['root'], temp=1, force_full_search=True, noise=case['noise'])
        assert (raw == case['raw_counts']).all(), str(case['cfg'])
        np.testing.assert_allclose(probs, case['probs'], rtol=0, atol=1e-12)
        assert (q == case['q']).all()
        st = m.stats()
        assert list(st[:3]) == list(case['summary'])


def test_episode_tree_reuse_exact(episodes):
    for tag, ep in episodes.items():
        cfg, noise = _cfg(str(ep['cfg']), ep['n_sims'])
        m = O.MCTS(cfg, dirichlet_noise=noise)
        for i in range(len(ep['roots'])):
            nz = ep['noise'][i][:ep['noise_len'][i]]
            probs, q, full, raw = m.getActionProb(ep['roots'][i], temp=1, force_full_search=True, noise=nz)
            assert (raw == ep['raw_counts'][i]).all(), f'{tag} ply {i}'
            np.testing.assert_allclose(probs, ep['probs'][i], rtol=0, atol=1e-12)
            assert (q == ep['q'][i]).all(), f'{tag} ply {i}'
            assert list(m.stats()[:3]) == list(ep['summaries'][i]), f'{tag} ply {i}'
