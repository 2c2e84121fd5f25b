"""ctypes front-end of the CPU oracle (oracle/azg_oracle.c). TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

NA = 81

# state_dict tensor order expected by azg_oracle.c:v80_bind (names as in splendor/SplendorNNet.py V80)
def _lin_bn(prefix):
    return [f'{prefix}.linear.weight', f'{prefix}.norm.weight', f'{prefix}.norm.bias',
            f'{prefix}.norm.running_mean', f'{prefix}.norm.running_var']


def v80_order():
    names = _lin_bn('first_layer')
    for blk in ('trunk.0', 'output_layers_PI.0', 'output_layers_V.0'):
        names += _lin_bn(f'{blk}.expand') + _lin_bn(f'{blk}.depthwise')
        names += [f'{blk}.se.fc1.weight', f'{blk}.se.fc1.bias', f'{blk}.se.fc2.weight', f'{blk}.se.fc2.bias']
        names += _lin_bn(f'{blk}.project')
    names += ['output_layers_PI.2.weight', 'output_layers_PI.2.bias', 'output_layers_PI.4.weight', 'output_layers_PI.4.bias',
              'output_layers_V.2.weight', 'output_layers_V.2.bias', 'output_layers_V.4.weight', 'output_layers_V.4.bias']
    return names


def v80_blob(state_dict):
    """Flatten a V80 state_dict (name -> ndarray) into the oracle's float32 blob."""
    return np.concatenate([np.asarray(state_dict[n], dtype=np.float32).ravel() for n in v80_order()]).astype(np.float32)


def build(force=False):
    so = os.path.join(HERE, 'libazg_oracle.so')
    src = os.path.join(HERE, 'azg_oracle.c')
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', HERE, '-s', '-B', 'libazg_oracle.so'])
    return so


class Cfg(C.Structure):
    _fields_ = [('num_players', C.c_int), ('numMCTSSims', C.c_int), ('ratio_fullMCTS', C.c_int), ('universes', C.c_int),
                ('forced_playouts', C.c_int), ('no_mem_optim', C.c_int), ('net_kind', C.c_int),
                ('cpuct', C.c_double), ('fpu', C.c_double), ('dirichletAlpha', C.c_double), ('prob_fullMCTS', C.c_double),
                ('temperature2', C.c_double), ('game', C.c_int)]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        p8, pu8, pf, pd, pi64 = (C.POINTER(C.c_int8), C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int64))
        L.azo_init_game.argtypes = [p8, C.c_int, C.c_uint64]
        L.azo_get_round.argtypes = [p8]; L.azo_get_round.restype = C.c_int
        L.azo_get_score.argtypes = [p8, C.c_int, C.c_int]; L.azo_get_score.restype = C.c_int
        L.azo_valid_moves.argtypes = [p8, C.c_int, C.c_int, pu8]
        L.azo_next_state.argtypes = [p8, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_uint64]; L.azo_next_state.restype = C.c_int
        L.azo_check_end_game.argtypes = [p8, C.c_int, pf]
        L.azo_swap_players.argtypes = [p8, C.c_int, C.c_int]
        L.azo_symmetries.argtypes = [p8, C.c_int, pf, pu8, p8, pf, pu8]; L.azo_symmetries.restype = C.c_int
        L.azo_hashnet.argtypes = [p8, C.c_int, pu8, C.c_int, pf, pf]
        L.azo_hashnet_a.argtypes = [p8, C.c_int, pu8, C.c_int, C.c_int, pf, pf]
        L.azo_sant_get_round.argtypes = [p8]; L.azo_sant_get_round.restype = C.c_int
        L.azo_sant_get_score.argtypes = [p8, C.c_int]; L.azo_sant_get_score.restype = C.c_int
        L.azo_sant_valid_moves.argtypes = [p8, C.c_int, pu8]
        L.azo_sant_make_move.argtypes = [p8, C.c_int, C.c_int]; L.azo_sant_make_move.restype = C.c_int
        L.azo_sant_check_end_game.argtypes = [p8, C.c_int, pf]
        L.azo_sant_swap_players.argtypes = [p8, C.c_int]
        L.azo_sant_init_game.argtypes = [p8, C.c_uint64]
        L.azo_sant_symmetries.argtypes = [p8, pf, pu8, p8, pf, pu8]; L.azo_sant_symmetries.restype = C.c_int
        L.azo_aba_get_round.argtypes = [p8]; L.azo_aba_get_round.restype = C.c_int
        L.azo_aba_get_score.argtypes = [p8, C.c_int]; L.azo_aba_get_score.restype = C.c_int
        L.azo_aba_valid_moves.argtypes = [p8, C.c_int, pu8]
        L.azo_aba_make_move.argtypes = [p8, C.c_int, C.c_int]; L.azo_aba_make_move.restype = C.c_int
        L.azo_aba_check_end_game.argtypes = [p8, pf]
        L.azo_aba_swap_players.argtypes = [p8, C.c_int]
        L.azo_aba_init_game.argtypes = [p8]
        L.azo_aba_symmetries.argtypes = [p8, pf, pu8, p8, pf, pu8]; L.azo_aba_symmetries.restype = C.c_int
        L.azo_azul_get_round.argtypes = [p8]; L.azo_azul_get_round.restype = C.c_int
        L.azo_azul_get_score.argtypes = [p8, C.c_int]; L.azo_azul_get_score.restype = C.c_int
        L.azo_azul_valid_moves.argtypes = [p8, C.c_int, pu8]
        L.azo_azul_make_move.argtypes = [p8, C.c_int, C.c_int, C.c_int64, C.c_uint64]; L.azo_azul_make_move.restype = C.c_int
        L.azo_azul_check_end_game.argtypes = [p8, pf]
        L.azo_azul_swap_players.argtypes = [p8]
        L.azo_azul_init_game.argtypes = [p8, C.c_uint64]
        L.azo_azul_symmetries.argtypes = [p8, pf, pu8, p8, pf, pu8]; L.azo_azul_symmetries.restype = C.c_int
        L.azo_v84_forward.argtypes = [pf, C.c_int, p8, pu8, pf, pf]
        L.azo_v80_forward.argtypes = [pf, C.c_int, C.c_int, p8, pu8, pf, pf]
        L.azo_v89_forward.argtypes = [pf, C.c_int, p8, pu8, pf, pf]
        L.azo_v21_forward.argtypes = [pf, C.c_int, p8, pu8, pf, pf]
        L.azo_mcts_new.argtypes = [C.POINTER(Cfg), pf, C.c_int, C.c_uint64]; L.azo_mcts_new.restype = C.c_void_p
        L.azo_mcts_free.argtypes = [C.c_void_p]
        L.azo_mcts_reset.argtypes = [C.c_void_p]
        L.azo_mcts_stats.argtypes = [C.c_void_p, pi64]
        L.azo_mcts_get_action_prob.argtypes = [C.c_void_p, p8, C.c_double, C.c_int, pd, pd, pf, pi64]; L.azo_mcts_get_action_prob.restype = C.c_int
        L.azo_selfplay_bench.argtypes = [C.POINTER(Cfg), pf, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_uint64, pd]
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _board(b):
    return np.ascontiguousarray(b, dtype=np.int8).copy()


def init_game(seed, n=2):
    b = np.zeros((32 + 10 * n + n * n, 7), np.int8)
    lib().azo_init_game(_p(b, C.c_int8), n, seed)
    return b


def valid_moves(board, player=0, n=2):
    b = _board(board); out = np.zeros(NA, np.uint8)
    lib().azo_valid_moves(_p(b, C.c_int8), n, player, _p(out, C.c_uint8))
    return out.astype(np.bool_)


def next_state(board, player, action, seed, n=2, rng_seed=0):
    b = _board(board)
    np_ = lib().azo_next_state(_p(b, C.c_int8), n, int(action), int(player), int(seed), int(rng_seed))
    return b, np_


def game_ended(board, n=2):
    b = _board(board); out = np.zeros(n, np.float32)
    lib().azo_check_end_game(_p(b, C.c_int8), n, _p(out, C.c_float))
    return out


def canonical(board, player, n=2):
    b = _board(board)
    if player:
        lib().azo_swap_players(_p(b, C.c_int8), n, int(player))
    return b


def get_round(board):
    b = _board(board)
    return lib().azo_get_round(_p(b, C.c_int8))


def get_score(board, player, n=2):
    b = _board(board)
    return lib().azo_get_score(_p(b, C.c_int8), n, player)


def symmetries(board, pi, valids, n=2):
    b = _board(board); pi = np.ascontiguousarray(pi, np.float32); v = np.ascontiguousarray(valids).astype(np.uint8)
    kmax = 1 + 9 + 2 * n
    ob = np.zeros((kmax,) + b.shape, np.int8); op = np.zeros((kmax, NA), np.float32); ov = np.zeros((kmax, NA), np.uint8)
    k = lib().azo_symmetries(_p(b, C.c_int8), n, _p(pi, C.c_float), _p(v, C.c_uint8), _p(ob, C.c_int8), _p(op, C.c_float), _p(ov, C.c_uint8))
    return [(ob[i], op[i], ov[i].astype(np.bool_)) for i in range(k)]


def hashnet(board, valids, n=2):
    b = _board(board); v = np.ascontiguousarray(valids).astype(np.uint8)
    pi = np.zeros(v.size, np.float32); val = np.zeros(n, np.float32)
    lib().azo_hashnet_a(_p(b, C.c_int8), b.size, _p(v, C.c_uint8), v.size, n, _p(pi, C.c_float), _p(val, C.c_float))
    return pi, val


def v80_forward(blob, boards, valids, n=2):
    boards = np.ascontiguousarray(boards, np.int8); B = boards.shape[0]
    v = np.ascontiguousarray(valids).astype(np.uint8)
    blob = np.ascontiguousarray(blob, np.float32)
    pi = np.zeros((B, NA), np.float32); val = np.zeros((B, n), np.float32)
    lib().azo_v80_forward(_p(blob, C.c_float), n, B, _p(boards, C.c_int8), _p(v, C.c_uint8), _p(pi, C.c_float), _p(val, C.c_float))
    return pi, val


GAME_SPLENDOR, GAME_SANTORINI, GAME_ABALONE, GAME_AZUL = 0, 1, 2, 3
GAME_ACTIONS = {GAME_SPLENDOR: 81, GAME_SANTORINI: 162, GAME_ABALONE: 3402, GAME_AZUL: 180}


def make_cfg(num_players=2, numMCTSSims=800, ratio_fullMCTS=5, universes=1, forced_playouts=False, no_mem_optim=False,
             net_kind=0, cpuct=1.25, fpu=0.0, dirichletAlpha=-1.0, prob_fullMCTS=1.0, temperature2=1.1, game=GAME_SPLENDOR):
    return Cfg(num_players, numMCTSSims, ratio_fullMCTS, universes, int(forced_playouts), int(no_mem_optim), net_kind,
               cpuct, fpu, dirichletAlpha, prob_fullMCTS, temperature2, game)


class MCTS:
    """Oracle counterpart of the reference's MCTS class (MCTS.py:19-203)."""

    def __init__(self, cfg, blob=None, dirichlet_noise=False, seed=0):
        self.cfg = cfg
        self.blob = None if blob is None else np.ascontiguousarray(blob, np.float32)
        bp = _p(self.blob, C.c_float) if self.blob is not None else None
        self.h = lib().azo_mcts_new(C.byref(cfg), bp, int(dirichlet_noise), seed)

    def __del__(self):
        if getattr(self, 'h', None):
            lib().azo_mcts_free(self.h); self.h = None

    def reset(self):
        lib().azo_mcts_reset(self.h)

    def stats(self):
        out = np.zeros(7, np.int64); lib().azo_mcts_stats(self.h, _p(out, C.c_int64)); return out

    def getActionProb(self, cb, temp=1.0, force_full_search=False, noise=None):
        A = GAME_ACTIONS[self.cfg.game]
        b = _board(cb); probs = np.zeros(A, np.float64); q = np.zeros(self.cfg.num_players, np.float32); raw = np.zeros(A, np.int64)
        nz = None
        if noise is not None and len(noise):
            nz = np.ascontiguousarray(noise, np.float64)
        full = lib().azo_mcts_get_action_prob(self.h, _p(b, C.c_int8), float(temp), int(force_full_search),
                                              _p(nz, C.c_double) if nz is not None else None,
                                              _p(probs, C.c_double), _p(q, C.c_float), _p(raw, C.c_int64))
        return probs, q, bool(full), raw


GAME_SHAPES = {GAME_SPLENDOR: (56, 7), GAME_SANTORINI: (5, 5, 3), GAME_ABALONE: (9, 9, 4), GAME_AZUL: (23, 6)}


def execute_episode_inj(cfg, init_board, u_full, u_move, chance_seed, noise=None, temperature=(1.0, 0.1), tempThreshold=10, blob=None,
                        dirichlet_noise=None):
    """Coach.executeEpisode (Coach.py:37-84) on the oracle with injected randomness (see azo_execute_episode_inj).
    noise: list of per-ply Dirichlet vectors (may be ragged / empty for non-full plies) or None.
    Returns dict(boards, pi, z, valids, q: un-augmented examples; plies; actions; full)."""
    A = GAME_ACTIONS[cfg.game]; npl = cfg.num_players
    shape = GAME_SHAPES[cfg.game] if cfg.game != GAME_SPLENDOR else (32 + 10 * npl + npl * npl, 7)
    P = len(u_full)
    dn = (cfg.dirichletAlpha != 0) if dirichlet_noise is None else dirichlet_noise
    m = MCTS(cfg, blob, dirichlet_noise=dn, seed=1)
    b0 = _board(init_board)
    uf = np.ascontiguousarray(u_full, np.float64); um = np.ascontiguousarray(u_move, np.float64)
    cs = np.ascontiguousarray(chance_seed if chance_seed is not None else np.ones(P), np.int64)
    nz = None; stride = 0
    if noise is not None:
        stride = max(max((len(x) for x in noise), default=0), 1)
        nz = np.zeros((P, stride), np.float64)
        for i, x in enumerate(noise):
            nz[i, :len(x)] = x
    eb = np.zeros((P,) + shape, np.int8); ep = np.zeros((P, A), np.float32); ez = np.zeros((P, npl), np.float32)
    ev = np.zeros((P, A), np.uint8); eq = np.zeros((P, npl), np.float32)
    plies = C.c_int(0); acts = np.zeros(P, np.int32); full = np.zeros(P, np.uint8)
    L = lib()
    L.azo_execute_episode_inj.restype = C.c_int
    n = L.azo_execute_episode_inj(C.c_void_p(m.h), _p(b0, C.c_int8), C.c_int(P), _p(uf, C.c_double), _p(um, C.c_double), _p(cs, C.c_int64),
                                  _p(nz, C.c_double) if nz is not None else None, C.c_int(stride), C.c_double(temperature[0]),
                                  C.c_double(temperature[1]), C.c_double(tempThreshold), C.c_int(P), _p(eb, C.c_int8), _p(ep, C.c_float),
                                  _p(ez, C.c_float), _p(ev, C.c_uint8), _p(eq, C.c_float), C.byref(plies), _p(acts, C.c_int32), _p(full, C.c_uint8))
    if n < 0:
        raise RuntimeError(f'episode did not finish within the {P} injected plies')
    k = plies.value
    return dict(boards=eb[:n], pi=ep[:n], z=ez[:n], valids=ev[:n].astype(np.bool_), q=eq[:n], plies=k, actions=acts[:k], full=full[:k].astype(np.bool_))


def symmetries_for(game):
    """getSymmetries of oracle game id `game` as f(board, pi, valids) -> [(board, pi, valids)]."""
    return {GAME_SPLENDOR: symmetries, GAME_SANTORINI: sant_symmetries, GAME_ABALONE: aba_symmetries, GAME_AZUL: azul_symmetries}[game]


def augment(game, ex):
    """Coach.py:66-69 applied to un-augmented examples: list of (board, pi, z, valids, q) in the reference's order."""
    sym = symmetries_for(game); out = []
    for i in range(len(ex['boards'])):
        for b, p, v in sym(ex['boards'][i], ex['pi'][i], ex['valids'][i]):
            out.append((b, p, ex['z'][i], v, ex['q'][i]))
    return out


def selfplay_bench(cfg, blob, threads, games_per_thread, max_plies=0, temperature=(1.0, 0.1), tempThreshold=10, seed=1):
    blob = np.ascontiguousarray(blob, np.float32); out = np.zeros(8, np.float64)
    lib().azo_selfplay_bench(C.byref(cfg), _p(blob, C.c_float), threads, games_per_thread, max_plies,
                             float(temperature[0]), float(temperature[1]), float(tempThreshold), seed, _p(out, C.c_double))
    keys = ('sims', 'expansions', 'node_visits', 'nn_evals', 'plies', 'examples', 'games', 'seconds')
    return dict(zip(keys, out.tolist()))


# ---- Santorini without gods (santorini/SantoriniLogicNumba.py with NB_GODS = 1) ----
SAN_A = 162


def sant_init_game(seed):
    b = np.zeros((5, 5, 3), np.int8); lib().azo_sant_init_game(_p(b, C.c_int8), seed); return b


def sant_valid_moves(board, player=0):
    b = _board(board); out = np.zeros(SAN_A, np.uint8)
    lib().azo_sant_valid_moves(_p(b, C.c_int8), int(player), _p(out, C.c_uint8))
    return out.astype(np.bool_)


def sant_next_state(board, player, action):
    b = _board(board)
    return b, lib().azo_sant_make_move(_p(b, C.c_int8), int(action), int(player))


def sant_game_ended(board, next_player):
    b = _board(board); out = np.zeros(2, np.float32)
    lib().azo_sant_check_end_game(_p(b, C.c_int8), int(next_player), _p(out, C.c_float))
    return out


def sant_canonical(board, player):
    b = _board(board)
    if player:
        lib().azo_sant_swap_players(_p(b, C.c_int8), int(player))
    return b


def sant_get_round(board):
    b = _board(board); return lib().azo_sant_get_round(_p(b, C.c_int8))


def sant_get_score(board, player):
    b = _board(board); return lib().azo_sant_get_score(_p(b, C.c_int8), int(player))


def sant_symmetries(board, pi, valids):
    b = _board(board); pi = np.ascontiguousarray(pi, np.float32); v = np.ascontiguousarray(valids).astype(np.uint8)
    ob = np.zeros((8, 5, 5, 3), np.int8); op = np.zeros((8, SAN_A), np.float32); ov = np.zeros((8, SAN_A), np.uint8)
    k = lib().azo_sant_symmetries(_p(b, C.c_int8), _p(pi, C.c_float), _p(v, C.c_uint8), _p(ob, C.c_int8), _p(op, C.c_float), _p(ov, C.c_uint8))
    return [(ob[i], op[i], ov[i].astype(np.bool_)) for i in range(k)]


def _bn(prefix):
    return [f'{prefix}.weight', f'{prefix}.bias', f'{prefix}.running_mean', f'{prefix}.running_var']


def v89_order():
    """state_dict tensor order expected by azg_oracle.c:v89_bind (names as in santorini/SantoriniNNet.py V89)."""
    names = ['first_layer.0.weight'] + _bn('first_layer.1')
    for b in range(5):
        names += [f'trunk.{b}.conv1.weight'] + _bn(f'trunk.{b}.bn1') + [f'trunk.{b}.conv2.weight'] + _bn(f'trunk.{b}.bn2')
    names += ['head_PI.conv1x1.weight'] + _bn('head_PI.bn') + ['head_PI.fc.weight', 'head_PI.fc.bias']
    names += ['head_V.conv1x1.weight'] + _bn('head_V.bn') + ['head_V.fc1.weight', 'head_V.fc1.bias', 'head_V.fc2.weight', 'head_V.fc2.bias']
    return names


def v89_blob(state_dict):
    return np.concatenate([np.asarray(state_dict[n], dtype=np.float32).ravel() for n in v89_order()]).astype(np.float32)


def v89_forward(blob, boards, valids):
    boards = np.ascontiguousarray(boards, np.int8); B = boards.shape[0]
    v = np.ascontiguousarray(valids).astype(np.uint8); blob = np.ascontiguousarray(blob, np.float32)
    pi = np.zeros((B, SAN_A), np.float32); val = np.zeros((B, 2), np.float32)
    lib().azo_v89_forward(_p(blob, C.c_float), B, _p(boards, C.c_int8), _p(v, C.c_uint8), _p(pi, C.c_float), _p(val, C.c_float))
    return pi, val


# ---- Abalone, Belgian daisy (abalone/AbaloneLogicNumba.py, shipped constants) ----
ABA_A = 3402


def aba_init_game():
    b = np.zeros((9, 9, 4), np.int8); lib().azo_aba_init_game(_p(b, C.c_int8)); return b


def aba_valid_moves(board, player=0):
    b = _board(board); out = np.zeros(ABA_A, np.uint8)
    lib().azo_aba_valid_moves(_p(b, C.c_int8), int(player), _p(out, C.c_uint8))
    return out.astype(np.bool_)


def aba_next_state(board, player, action):
    b = _board(board)
    return b, lib().azo_aba_make_move(_p(b, C.c_int8), int(action), int(player))


def aba_game_ended(board):
    b = _board(board); out = np.zeros(2, np.float32)
    lib().azo_aba_check_end_game(_p(b, C.c_int8), _p(out, C.c_float))
    return out


def aba_canonical(board, player):
    b = _board(board)
    if player:
        lib().azo_aba_swap_players(_p(b, C.c_int8), int(player))
    return b


def aba_get_round(board):
    b = _board(board); return lib().azo_aba_get_round(_p(b, C.c_int8))


def aba_get_score(board, player):
    b = _board(board); return lib().azo_aba_get_score(_p(b, C.c_int8), int(player))


def aba_symmetries(board, pi, valids):
    b = _board(board); pi = np.ascontiguousarray(pi, np.float32); v = np.ascontiguousarray(valids).astype(np.uint8)
    ob = np.zeros((12, 9, 9, 4), np.int8); op = np.zeros((12, ABA_A), np.float32); ov = np.zeros((12, ABA_A), np.uint8)
    k = lib().azo_aba_symmetries(_p(b, C.c_int8), _p(pi, C.c_float), _p(v, C.c_uint8), _p(ob, C.c_int8), _p(op, C.c_float), _p(ov, C.c_uint8))
    return [(ob[i], op[i], ov[i].astype(np.bool_)) for i in range(k)]


# ---- Azul, 2 players (azul/AzulLogicNumba.py) -- rules + MCTS (GAME_AZUL), round-2 groundwork ----
AZUL_A = 180


def azul_init_game(seed=0):
    b = np.zeros((23, 6), np.int8); lib().azo_azul_init_game(_p(b, C.c_int8), int(seed)); return b


def azul_valid_moves(board, player=0):
    b = _board(board); out = np.zeros(AZUL_A, np.uint8)
    lib().azo_azul_valid_moves(_p(b, C.c_int8), int(player), _p(out, C.c_uint8))
    return out.astype(np.bool_)


def azul_next_state(board, player, action, random_seed, rng_seed=0):
    b = _board(board)
    return b, lib().azo_azul_make_move(_p(b, C.c_int8), int(action), int(player), int(random_seed), int(rng_seed))


def azul_game_ended(board):
    b = _board(board); out = np.zeros(2, np.float32)
    lib().azo_azul_check_end_game(_p(b, C.c_int8), _p(out, C.c_float))
    return out


def azul_canonical(board, player):
    b = _board(board)
    if player:
        lib().azo_azul_swap_players(_p(b, C.c_int8))
    return b


def azul_get_round(board):
    b = _board(board); return lib().azo_azul_get_round(_p(b, C.c_int8))


def azul_get_score(board, player):
    b = _board(board); return lib().azo_azul_get_score(_p(b, C.c_int8), int(player))


def azul_symmetries(board, pi, valids):
    b = _board(board); pi = np.ascontiguousarray(pi, np.float32); v = np.ascontiguousarray(valids).astype(np.uint8)
    ob = np.zeros((120, 23, 6), np.int8); op = np.zeros((120, AZUL_A), np.float32); ov = np.zeros((120, AZUL_A), np.uint8)
    k = lib().azo_azul_symmetries(_p(b, C.c_int8), _p(pi, C.c_float), _p(v, C.c_uint8), _p(ob, C.c_int8), _p(op, C.c_float), _p(ov, C.c_uint8))
    return [(ob[i], op[i], ov[i].astype(np.bool_)) for i in range(k)]


def v84_blob(state_dict):
    """AzulNNet V84 (azul/AzulNNet.py:84-111) uses SplendorNNet V80's module names: same tensor order, different shapes."""
    sizes = {'first_layer.linear.weight': (23, 23), 'trunk.0.expand.linear.weight': (115, 23), 'trunk.0.se.fc1.weight': (32, 115),
             'output_layers_PI.0.project.linear.weight': (46, 115), 'output_layers_V.0.se.fc1.weight': (16, 46),
             'output_layers_PI.2.weight': (180, 276), 'output_layers_V.2.weight': (2, 138)}
    for k, shp in sizes.items():
        assert tuple(np.asarray(state_dict[k]).shape) == shp, (k, np.asarray(state_dict[k]).shape)
    return v80_blob(state_dict)


def v84_forward(blob, boards, valids):
    boards = np.ascontiguousarray(boards, np.int8); B = boards.shape[0]
    v = np.ascontiguousarray(valids).astype(np.uint8); blob = np.ascontiguousarray(blob, np.float32)
    pi = np.zeros((B, AZUL_A), np.float32); val = np.zeros((B, 2), np.float32)
    lib().azo_v84_forward(_p(blob, C.c_float), B, _p(boards, C.c_int8), _p(v, C.c_uint8), _p(pi, C.c_float), _p(val, C.c_float))
    return pi, val


def v21_order():
    """state_dict tensor order expected by azg_oracle.c:v21_bind (names as in abalone/AbaloneNNet.py V21)."""
    names = ['first_layer.0.weight'] + _bn('first_layer.1')
    for b in range(4):
        for j in range(3):
            names += [f'trunk.{b}.block.{j}.0.weight'] + _bn(f'trunk.{b}.block.{j}.1')
    names += ['meta_fc.0.weight', 'meta_fc.0.bias', 'head_PI.0.weight'] + _bn('head_PI.1') + ['head_V_conv.0.weight'] + _bn('head_V_conv.1')
    names += ['head_V_fc.0.weight', 'head_V_fc.0.bias', 'head_V_fc.2.weight', 'head_V_fc.2.bias']
    return names


def v21_blob(state_dict):
    return np.concatenate([np.asarray(state_dict[n], dtype=np.float32).ravel() for n in v21_order()]).astype(np.float32)


def v21_forward(blob, boards, valids):
    boards = np.ascontiguousarray(boards, np.int8); B = boards.shape[0]
    v = np.ascontiguousarray(valids).astype(np.uint8); blob = np.ascontiguousarray(blob, np.float32)
    pi = np.zeros((B, ABA_A), np.float32); val = np.zeros((B, 2), np.float32)
    lib().azo_v21_forward(_p(blob, C.c_float), B, _p(boards, C.c_int8), _p(v, C.c_uint8), _p(pi, C.c_float), _p(val, C.c_float))
    return pi, val
