import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def kat():
    return np.load(os.path.join(GOLDEN, 'splendor_kat.npz'))


@pytest.fixture(scope='session')
def mcts_cases():
    z = np.load(os.path.join(GOLDEN, 'splendor_mcts.npz'))
    n = int(z['n_cases'])
    keys = ('cfg', 'root', 'n_sims', 'probs', 'q', 'raw_counts', 'root_P', 'root_Qsa', 'noise', 'summary')
    return [{k: z[f'c{i}_{k}'] for k in keys} for i in range(n)]


@pytest.fixture(scope='session')
def episodes():
    return {t: np.load(os.path.join(GOLDEN, f'splendor_episode_{t}.npz')) for t in ('A', 'B')}


@pytest.fixture(scope='session')
def v80_golden():
    out = {}
    for tag in ('rand', 'shipped'):
        z = np.load(os.path.join(GOLDEN, f'splendor_v80_{tag}.npz'))
        sd = {k[4:]: z[k] for k in z.files if k.startswith('sd__')}
        out[tag] = dict(sd=sd, boards=z['boards'], valids=z['valids'], pi=z['pi'], v=z['v'])
    return out


# MCTS argument sets used to produce the goldens (mirrors oracle/gen_golden.py:MCTS_CONFIGS)
MCTS_CONFIGS = {
    'default': dict(cpuct=1.25, fpu=0.0, universes=1, dirichletAlpha=-1.0, temperature=[1.0, 0.1, 1.1],
                    forced_playouts=False, noise=False),
    'shipped': dict(cpuct=0.8, fpu=0.0593, universes=3, dirichletAlpha=0.3, temperature=[1.25, 0.8, 1.1],
                    forced_playouts=True, noise=True),
    'auto_noise': dict(cpuct=1.25, fpu=0.2, universes=0, dirichletAlpha=-1.0, temperature=[1.0, 0.1, 1.1],
                       forced_playouts=False, noise=True),
    'universes8': dict(cpuct=2.0, fpu=0.0, universes=8, dirichletAlpha=-1.0, temperature=[1.0, 0.1, 1.0],
                       forced_playouts=True, noise=False),
}


@pytest.fixture(scope='session')
def san_kat():
    return np.load(os.path.join(GOLDEN, 'santorini_kat.npz'))


@pytest.fixture(scope='session')
def san_mcts_cases():
    z = np.load(os.path.join(GOLDEN, 'santorini_mcts.npz'))
    n = int(z['n_cases'])
    keys = ('cfg', 'root', 'n_sims', 'probs', 'q', 'raw_counts', 'root_P', 'root_Qsa', 'noise', 'summary')
    return [{k: z[f'c{i}_{k}'] for k in keys} for i in range(n)]


@pytest.fixture(scope='session')
def san_episode():
    return np.load(os.path.join(GOLDEN, 'santorini_episode.npz'))


@pytest.fixture(scope='session')
def v89_golden():
    out = {}
    for tag in ('rand', 'shipped'):
        z = np.load(os.path.join(GOLDEN, f'santorini_v89_{tag}.npz'))
        sd = {k[4:]: z[k] for k in z.files if k.startswith('sd__')}
        out[tag] = dict(sd=sd, boards=z['boards'], valids=z['valids'], pi=z['pi'], v=z['v'])
    return out


def _unpack(bits, n=3402):
    return np.unpackbits(bits, axis=-1)[..., :n].astype(np.bool_)


@pytest.fixture(scope='session')
def aba_kat():
    z = np.load(os.path.join(GOLDEN, 'abalone_kat.npz'))
    d = {k: z[k] for k in z.files}
    d['valids'] = _unpack(d['valids']); d['sym_valids'] = _unpack(d['sym_valids']); d['sym_out_valids'] = _unpack(d['sym_out_valids'])
    return d


@pytest.fixture(scope='session')
def azul_kat():
    z = np.load(os.path.join(GOLDEN, 'azul_kat.npz'))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope='session')
def azul_mcts_cases():
    z = np.load(os.path.join(GOLDEN, 'azul_mcts.npz'))
    keys = ('cfg', 'root', 'n_sims', 'probs', 'q', 'raw_counts', 'noise', 'summary')
    return [{k: z[f'c{i}_{k}'] for k in keys} for i in range(int(z['n_cases']))]


@pytest.fixture(scope='session')
def azul_episode():
    z = np.load(os.path.join(GOLDEN, 'azul_episode.npz'))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope='session')
def aba_mcts_cases():
    return _aba_cases('abalone_mcts.npz')


@pytest.fixture(scope='session')
def aba_mcts1600_cases():
    """numMCTSSims = 1600, the length BASELINE.json configs[4] runs (oracle/gen_golden_abalone.py --only mcts1600)."""
    return _aba_cases('abalone_mcts1600.npz')


def _aba_cases(fname):
    z = np.load(os.path.join(GOLDEN, fname))
    n = int(z['n_cases'])
    keys = ('cfg', 'root', 'n_sims', 'q', 'raw_idx', 'raw_cnt', 'probs_nz', 'probs_idx', 'noise', 'summary')
    out = []
    for i in range(n):
        c = {k: z[f'c{i}_{k}'] for k in keys}
        raw = np.zeros(3402, np.int64); raw[c['raw_idx']] = c['raw_cnt']; c['raw_counts'] = raw
        probs = np.zeros(3402, np.float64); probs[c['probs_idx']] = c['probs_nz']; c['probs'] = probs
        out.append(c)
    return out


@pytest.fixture(scope='session')
def aba_episode():
    z = np.load(os.path.join(GOLDEN, 'abalone_episode.npz'))
    d = {k: z[k] for k in z.files}
    raw = np.zeros((len(d['roots']), 3402), np.int64)
    for i in range(len(raw)):
        m = d['raw_idx'][i] >= 0
        raw[i, d['raw_idx'][i][m]] = d['raw_cnt'][i][m]
    d['raw_counts'] = raw
    return d


@pytest.fixture(scope='session')
def v21_golden():
    out = {}
    for tag in ('rand', 'shipped'):
        z = np.load(os.path.join(GOLDEN, f'abalone_v21_{tag}.npz'))
        sd = {k[4:]: z[k] for k in z.files if k.startswith('sd__')}
        out[tag] = dict(sd=sd, boards=z['boards'], valids=_unpack(z['valids']), pi=z['pi'], v=z['v'])
    return out


@pytest.fixture(scope='session')
def v84_golden():
    out = {}
    for tag in ('rand', 'shipped'):
        z = np.load(os.path.join(GOLDEN, f'azul_v84_{tag}.npz'))
        sd = {k[4:]: z[k] for k in z.files if k.startswith('sd__')}
        out[tag] = dict(sd=sd, boards=z['boards'], valids=z['valids'], pi=z['pi'], v=z['v'])
    return out


def load_selfplay_golden(game):
    """tests/golden/<game>_selfplay.npz (oracle/gen_golden_selfplay.py): the reference's Coach.executeEpisode with every random
    input recorded. Returns (cfg dict, [game dicts with init, u_full, u_move, chance_seed, noise (ragged list), is_full, action,
    ex_board, ex_pi, ex_z, ex_valids, ex_q])."""
    z = np.load(os.path.join(GOLDEN, f'{game}_selfplay.npz'))
    cfg = {k[4:]: z[k].tolist() for k in z.files if k.startswith('cfg_')}
    games = []
    for gi in range(int(z['n_games'])):
        p = f'g{gi}_'
        d = {k: z[p + k] for k in ('init', 'u_full', 'u_move', 'chance_seed', 'is_full', 'action', 'root', 'ex_board', 'ex_z', 'ex_q')}
        d['noise'] = [z[p + 'noise'][i, :int(n)] for i, n in enumerate(z[p + 'noise_len'])]
        n_ex = len(d['ex_board'])
        if p + 'ex_pi' in z.files:
            d['ex_pi'] = z[p + 'ex_pi']; d['ex_valids'] = z[p + 'ex_valids']
        else:
            A = 3402
            pi = np.zeros((n_ex, A), np.float32); idx = z[p + 'ex_pi_idx']; pi[idx[0], idx[1]] = z[p + 'ex_pi_val']
            d['ex_pi'] = pi; d['ex_valids'] = np.unpackbits(z[p + 'ex_valids_bits'], axis=1)[:, :A].astype(np.bool_)
        games.append(d)
    return cfg, games


def assert_examples_equal(got, gold, pi_tol=0.0):
    """got: list of (board, pi, z, valids, q) tuples; gold: one game dict of load_selfplay_golden. Example for example."""
    assert len(got) == len(gold['ex_board']), (len(got), len(gold['ex_board']))
    for i, (b, pi, zz, v, q) in enumerate(got):
        assert (np.asarray(b).reshape(-1) == gold['ex_board'][i].reshape(-1)).all(), f'example {i}: board'
        if pi_tol == 0.0:
            assert (np.asarray(pi, np.float32) == gold['ex_pi'][i]).all(), f'example {i}: pi'
        else:
            assert np.abs(np.asarray(pi, np.float32) - gold['ex_pi'][i]).max() <= pi_tol, f'example {i}: pi'
        assert (np.asarray(zz, np.float32) == gold['ex_z'][i]).all(), f'example {i}: z {zz} vs {gold["ex_z"][i]}'
        assert (np.asarray(v).astype(bool) == gold['ex_valids'][i]).all(), f'example {i}: valids'
        assert (np.asarray(q, np.float32) == gold['ex_q'][i]).all(), f'example {i}: q'


def load_splendor_np(n):
    """Goldens of Splendor with n = 3 / 4 players (oracle/gen_golden_splendor_np.py): kat dict, MCTS cases, shipped-net vectors."""
    kat = dict(np.load(os.path.join(GOLDEN, f'splendor{n}p_kat.npz')))
    z = np.load(os.path.join(GOLDEN, f'splendor{n}p_mcts.npz'))
    keys = ('cfg', 'root', 'n_sims', 'probs', 'q', 'raw_counts', 'noise', 'summary')
    cases = [{k: z[f'c{i}_{k}'] for k in keys} for i in range(int(z['n_cases']))]
    zn = np.load(os.path.join(GOLDEN, f'splendor{n}p_v80_shipped.npz'))
    net = dict(sd={k[4:]: zn[k] for k in zn.files if k.startswith('sd__')}, boards=zn['boards'], valids=zn['valids'], pi=zn['pi'], v=zn['v'])
    return kat, cases, net
