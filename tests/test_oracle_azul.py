"""Pins the CPU oracle's Azul (2 players) rules to vectors produced by the UNMODIFIED reference (tests/golden/azul_kat.npz, made by
oracle/gen_golden_azul.py: 12 random games played to the end with `random_seed != 0`, i.e. the reference's deterministic tile draws).
azul_mcts.npz / azul_episode.npz: the reference's MCTS on Azul positions with the hash-net). Bit-exact.
The CUDA plugin (csrc/azul.cuh) is checked through the C ABI by tests/test_gpu_azul.py; its lane functions also run on the host here."""
import numpy as np

from conftest import MCTS_CONFIGS
from oracle import oracle as O


def test_valid_moves_bit_exact(azul_kat):
    k = azul_kat
    for i in range(len(k['action'])):
        assert (O.azul_valid_moves(k['canonical'][i], 0) == k['valids'][i]).all(), i
        assert (O.azul_valid_moves(k['board'][i], int(k['player'][i])) == k['valids'][i]).all(), i


def test_next_state_ended_round_score_canonical(azul_kat):
    k = azul_kat
    new_rounds = 0
    for i in range(len(k['action'])):
        nb, npl = O.azul_next_state(k['board'][i], k['player'][i], k['action'][i], k['seed'][i])
        assert npl == k['next_player'][i] and (nb == k['next_board'][i]).all(), f'ply {i} action {k["action"][i]}'
        assert (O.azul_game_ended(nb) == k['ended'][i]).all()
        assert O.azul_get_round(nb) == k['round'][i] and [O.azul_get_score(nb, 0), O.azul_get_score(nb, 1)] == list(k['score'][i])
        assert (O.azul_canonical(k['board'][i], k['player'][i]) == k['canonical'][i]).all()
        assert (O.azul_canonical(nb, npl) == k['next_canonical'][i]).all()
        new_rounds += int(O.azul_get_round(nb) != O.azul_get_round(k['board'][i]))
    # the fixture covers round scoring + refills (incl. bag refills from the discards), end-of-game bonuses and every game result
    assert new_rounds >= 60 and (np.abs(k['ended']).sum(1) > 0).sum() == 12
    assert (k['next_board'][:, 1, :5].sum(1) < 20).any()


def test_symmetries_all_120_factory_orders(azul_kat):
    k = azul_kat
    for i in range(len(k['sym_pi'])):
        s = O.azul_symmetries(k['sym_board'][i], k['sym_pi'][i], k['sym_valids'][i])
        assert len(s) == 120
        for j, (b, p, v) in enumerate(s):
            assert (b == k['sym_out_boards'][i][j]).all(), (i, j)
            assert (p == k['sym_out_pi'][i][j]).all(), (i, j)
            assert (v == k['sym_out_valids'][i][j]).all(), (i, j)


def test_init_game_invariants():
    for seed in range(20):
        b = O.azul_init_game(seed)
        assert b[1, :5].sum() == 80 and (b[4:9, :5].sum(1) == 4).all() and b[3, 5] == 1 and b[0, 2] == 1
        assert (b[9:11, :5] == -1).all() and (b[11:23] == 0).all() and O.azul_valid_moves(b, 0).sum() > 0


def _cfg(name, n_sims):
    c = MCTS_CONFIGS[name]
    return O.make_cfg(numMCTSSims=int(n_sims), universes=c['universes'], forced_playouts=c['forced_playouts'], cpuct=c['cpuct'], fpu=c['fpu'],
                      dirichletAlpha=c['dirichletAlpha'], temperature2=c['temperature'][2], net_kind=0, game=O.GAME_AZUL), c['noise']


def test_mcts_counts_exact(azul_mcts_cases):
    """Visit counts, policy, q and tree size of the reference's search (universes 0/1/3/8, forced playouts, injected Dirichlet noise)."""
    assert len(azul_mcts_cases) == 15
    for case in azul_mcts_cases:
        cfg, noise = _cfg(str(case['cfg']), case['n_sims'])
        m = O.MCTS(cfg, dirichlet_noise=noise)
        probs, q, full, raw = m.getActionProb(case['root'], temp=1, force_full_search=True, noise=case['noise'])
        assert (raw == case['raw_counts']).all(), str(case['cfg'])
        np.testing.assert_allclose(probs, case['probs'], rtol=0, atol=1e-12)
        assert (q == case['q']).all()
        assert list(m.stats()[:3]) == list(case['summary'])


def test_episode_tree_reuse_exact(azul_episode):
    ep = azul_episode
    cfg, _ = _cfg('default', ep['n_sims'])
    m = O.MCTS(cfg, dirichlet_noise=False)
    for i in range(len(ep['roots'])):
        probs, q, full, raw = m.getActionProb(ep['roots'][i], temp=1, force_full_search=True)
        assert (raw == ep['raw_counts'][i]).all(), f'ply {i}'
        assert (q == ep['q'][i]).all(), f'ply {i}'
        assert list(m.stats()[:3]) == list(ep['summaries'][i]), f'ply {i}'


import pytest


@pytest.mark.parametrize('tag', ['rand', 'shipped'])
def test_v84_forward_vs_reference_torch(v84_golden, tag):
    """AzulNNet V84 restatement vs the reference's torch CPU fp32 outputs (random-init with perturbed BN, and azul/pretrained.pt);
    tolerance 1e-5 absolute on pi and v, the same bar as the three built nets."""
    g = v84_golden[tag]
    pi, v = O.v84_forward(O.v84_blob(g['sd']), g['boards'], g['valids'])
    np.testing.assert_allclose(pi, g['pi'], rtol=0, atol=1e-5)
    np.testing.assert_allclose(v, g['v'], rtol=0, atol=1e-5)
    assert (pi[~g['valids']] == 0).all() and np.abs(pi.sum(1) - 1).max() < 1e-5


# ---- the device plugin (csrc/azul.cuh) run on the HOST: tests/host/azul_plugin_emul.cpp defines the CUDA
# ---- qualifiers away and runs the lanes of a warp function one after the other. Not the product path, not a GPU result.
@pytest.fixture(scope='module')
def azul_emul(tmp_path_factory):
    import ctypes as C
    import os
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gxx = shutil.which('g++')
    if gxx is None:
        pytest.skip('g++ not available')
    so = str(tmp_path_factory.mktemp('azul_emul') / 'libazul_emul.so')
    subprocess.run([gxx, '-std=c++17', '-O1', '-fPIC', '-shared', '-Wno-unknown-pragmas', '-o', so, os.path.join(root, 'tests', 'host', 'azul_plugin_emul.cpp')],
                   check=True, timeout=300)
    L = C.CDLL(so)
    L.emul_make_move.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_longlong]
    L.emul_valid.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.emul_ended.argtypes = [C.c_void_p, C.c_void_p]
    L.emul_swap.argtypes = [C.c_void_p, C.c_int]
    L.emul_round.argtypes = [C.c_void_p]; L.emul_score.argtypes = [C.c_void_p, C.c_int]
    L.emul_symmetries.argtypes = [C.c_void_p] * 6
    L.emul_init.argtypes = [C.c_void_p, C.c_uint64]
    return L


def _ptr(a):
    return a.ctypes.data


def test_device_plugin_rules_on_host(azul_kat, azul_emul):
    L, k = azul_emul, azul_kat
    sizes = np.zeros(5, np.int32); L.emul_sizes(_ptr(sizes))
    assert sizes.tolist() == [138, 144, 180, 120, 2]
    for i in range(len(k['action'])):
        board = np.ascontiguousarray(k['board'][i], np.int8); player = int(k['player'][i])
        v = np.zeros(180, np.uint8); L.emul_valid(_ptr(board), player, _ptr(v))
        assert (v.astype(bool) == k['valids'][i]).all(), i
        cb = board.copy(); L.emul_swap(_ptr(cb), player)
        assert (cb == k['canonical'][i]).all(), i
        v0 = np.zeros(180, np.uint8); L.emul_valid(_ptr(cb), 0, _ptr(v0))
        assert (v0 == v).all(), i
        nb = board.copy(); npl = L.emul_make_move(_ptr(nb), int(k['action'][i]), player, int(k['seed'][i]))
        assert npl == k['next_player'][i] and (nb == k['next_board'][i]).all(), f'ply {i} action {k["action"][i]}'
        es = np.zeros(2, np.float32); over = L.emul_ended(_ptr(nb), _ptr(es))
        assert (es == k['ended'][i]).all() and bool(over) == bool(np.abs(k['ended'][i]).sum() > 0)
        assert L.emul_round(_ptr(nb)) == k['round'][i] and [L.emul_score(_ptr(nb), 0), L.emul_score(_ptr(nb), 1)] == list(k['score'][i])
        ncb = nb.copy(); L.emul_swap(_ptr(ncb), int(npl))
        assert (ncb == k['next_canonical'][i]).all(), i


def test_device_plugin_symmetries_on_host(azul_kat, azul_emul):
    L, k = azul_emul, azul_kat
    for i in range(len(k['sym_pi'])):
        b = np.ascontiguousarray(k['sym_board'][i], np.int8); pi = np.ascontiguousarray(k['sym_pi'][i], np.float32)
        m = np.ascontiguousarray(k['sym_valids'][i]).astype(np.uint8)
        ob = np.zeros((120, 23, 6), np.int8); op = np.zeros((120, 180), np.float32); om = np.zeros((120, 180), np.uint8)
        assert L.emul_symmetries(_ptr(b), _ptr(pi), _ptr(m), _ptr(ob), _ptr(op), _ptr(om)) == 120
        assert (ob == k['sym_out_boards'][i]).all() and (op == k['sym_out_pi'][i]).all() and (om.astype(bool) == k['sym_out_valids'][i]).all(), i
    for seed in range(8):
        b = np.zeros((23, 6), np.int8); L.emul_init(_ptr(b), seed)
        assert b[1, :5].sum() == 80 and (b[4:9, :5].sum(1) == 4).all() and b[3, 5] == 1 and b[0, 2] == 1 and (b[9:11, :5] == -1).all()
