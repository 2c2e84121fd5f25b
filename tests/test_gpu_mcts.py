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
