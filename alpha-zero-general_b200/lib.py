"""ctypes binding of include/azg.h (csrc/libazg_b200.so). Fails loudly if the library is missing."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
SO_PATH = os.environ.get('AZG_LIB_PATH') or os.path.join(CSRC, 'libazg_b200.so')   # AZG_LIB_PATH: another build of the same ABI (A/B timing runs)

AZG_GAME_SPLENDOR = 1
AZG_GAME_SANTORINI = 2
AZG_GAME_ABALONE = 3
AZG_GAME_AZUL = 4
AZG_ABI_VERSION = 4
AZG_N_STATS = 20
AZG_NET_HASH = 0
AZG_NET_SPLENDOR_V80 = 80
AZG_NET_SANTORINI_V89 = 89
AZG_NET_ABALONE_V21 = 21
AZG_NET_AZUL_V84 = 84

# every symbol include/azg.h declares (checked by tests/test_abi.py)
SYMBOLS = ['azg_abi_version', 'azg_last_error', 'azg_device_count', 'azg_set_device', 'azg_engine_profile', 'azg_engine_kernel_times', 'azg_game_info', 'azg_game_init', 'azg_game_valid',
           'azg_game_next', 'azg_game_ended', 'azg_game_canonical', 'azg_game_round_score', 'azg_game_symmetries',
           'azg_net_create', 'azg_net_load', 'azg_net_forward', 'azg_net_destroy', 'azg_engine_create', 'azg_engine_destroy',
           'azg_engine_reset', 'azg_engine_search', 'azg_engine_selfplay', 'azg_engine_selfplay_inject', 'azg_engine_selfplay_state', 'azg_engine_node', 'azg_engine_examples', 'azg_engine_examples_pending', 'azg_engine_stats', 'azg_net_prof', 'azg_net_prof_ctas', 'azg_debug_selprof']


class GameInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('game_id', 'num_players', 'state_rows', 'state_cols', 'state_depth', 'state_bytes', 'action_size',
                                         'max_symmetries', 'max_game_len')]


class EngineCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('game_id', 'num_players', 'n_games', 'numMCTSSims', 'ratio_fullMCTS', 'universes',
                                         'forced_playouts', 'no_mem_optim', 'dirichlet_noise', 'node_cap', 'edge_cap')] + \
               [(n, C.c_double) for n in ('cpuct', 'fpu', 'dirichletAlpha', 'prob_fullMCTS')] + \
               [('temperature', C.c_double * 3), ('tempThreshold', C.c_double), ('seed', C.c_uint64), ('first_game', C.c_uint64)]


class SelfplayInject(C.Structure):
    _fields_ = [('n_plies', C.c_int32), ('init_boards', C.c_void_p), ('u_full', C.c_void_p), ('u_move', C.c_void_p), ('chance_seed', C.c_void_p),
                ('noise', C.c_void_p)]


class AzgError(RuntimeError):
    pass


def build(verbose=False):
    """Compile csrc/ for sm_100a with nvcc (cross-compiles without a GPU)."""
    subprocess.check_call(['make', '-C', CSRC] + ([] if verbose else ['-s']))
    return SO_PATH


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise AzgError(f'{SO_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                       '(make -C alpha-zero-general_b200/csrc). There is no CPU fallback.')
    L = C.CDLL(SO_PATH)
    vp, i32, u64, sz = C.c_void_p, C.c_int32, C.c_uint64, C.c_size_t
    L.azg_abi_version.restype = i32
    L.azg_last_error.restype = C.c_char_p
    L.azg_device_count.restype = i32
    L.azg_game_info.argtypes = [i32, i32, C.POINTER(GameInfo)]
    L.azg_game_init.argtypes = [i32, i32, i32, vp, vp, vp]
    L.azg_game_valid.argtypes = [i32, i32, i32, vp, vp, vp, vp]
    L.azg_game_next.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.azg_game_ended.argtypes = [i32, i32, i32, vp, vp, vp, vp]
    L.azg_game_canonical.argtypes = [i32, i32, i32, vp, vp, vp, vp]
    L.azg_game_round_score.argtypes = [i32, i32, i32, vp, vp, vp, vp]
    L.azg_game_symmetries.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.azg_net_create.argtypes = [i32, i32, i32, vp, sz, C.POINTER(vp)]
    L.azg_net_load.argtypes = [vp, vp, sz]
    L.azg_net_forward.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.azg_net_destroy.argtypes = [vp]
    L.azg_engine_create.argtypes = [C.POINTER(EngineCfg), vp, C.POINTER(vp)]
    L.azg_engine_destroy.argtypes = [vp]
    L.azg_engine_reset.argtypes = [vp, i32]
    L.azg_engine_search.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp]
    L.azg_engine_selfplay.argtypes = [vp, i32, i32, vp]
    L.azg_engine_selfplay_inject.argtypes = [vp, C.POINTER(SelfplayInject)]
    L.azg_engine_selfplay_state.argtypes = [vp, vp, vp, vp, vp]
    L.azg_engine_node.argtypes = [vp, i32] + [vp] * 12
    L.azg_engine_examples.argtypes = [vp, i32, vp, vp, vp, vp, vp, C.POINTER(i32)]
    L.azg_engine_examples_pending.argtypes = [vp, C.POINTER(i32)]
    L.azg_engine_stats.argtypes = [vp, vp]
    L.azg_set_device.argtypes = [i32]
    L.azg_engine_profile.argtypes = [vp, i32]
    L.azg_engine_kernel_times.argtypes = [vp, vp]
    for name in SYMBOLS:
        if name not in ('azg_last_error',):
            getattr(L, name).restype = i32
    if L.azg_abi_version() != AZG_ABI_VERSION:
        raise AzgError(f'{SO_PATH}: ABI version {L.azg_abi_version()} != {AZG_ABI_VERSION}; rebuild (make -C alpha-zero-general_b200/csrc)')
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise AzgError(load().azg_last_error().decode())


def ptr(x):
    """Device or host pointer of a numpy array / torch tensor / None, as an int for ctypes c_void_p."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        assert x.flags['C_CONTIGUOUS'], 'array must be C-contiguous'
        return x.ctypes.data
    if hasattr(x, 'data_ptr'):
        assert x.is_contiguous(), 'tensor must be contiguous'
        return x.data_ptr()
    raise TypeError(type(x))


def device_count():
    return int(load().azg_device_count())


def game_info(game_id=AZG_GAME_SPLENDOR, num_players=2):
    gi = GameInfo()
    check(load().azg_game_info(game_id, num_players, C.byref(gi)))
    return gi


STAT_NAMES = ['sims', 'node_visits', 'expansions', 'nn_evals', 'terminal_hits', 'arena_overflows', 'gc_runs', 'max_nodes',
              'sum_legal', 'moves_played', 'episodes_finished', 'examples_recorded', 'kernels_launched', 'gc_sweeps', 'node_cap',
              'sum_legal_visited', 'sum_legal_root_scans', 'sum_legal_refreshed', 'examples_dropped', 'gc_trims']
