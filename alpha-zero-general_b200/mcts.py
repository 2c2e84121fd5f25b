"""MCTS facade (reference: MCTS.py:19-203) over the batched search engine.

`MCTS(game, nnet, args, dirichlet_noise=False)` + `getActionProb(canonicalBoard, temp, force_full_search)` keep the
reference's signature for one tree; `Engine` exposes the batched form the self-play loop uses.
"""
import ctypes as C

import numpy as np

from . import lib as _lib
from .utils import with_defaults


class Engine:
    """n_games concurrent trees on one GPU (azg_engine_*)."""

    def __init__(self, game, nnet, args, n_games, dirichlet_noise=False, seed=0, node_cap=0, edge_cap=0):
        self._L = _lib.load()
        self.game = game; self.nnet = nnet; self.args = a = with_defaults(args); self.n_games = n_games
        cfg = _lib.EngineCfg()
        cfg.game_id = game.game_id; cfg.num_players = game.num_players; cfg.n_games = n_games
        cfg.numMCTSSims = int(a.numMCTSSims); cfg.ratio_fullMCTS = int(a.ratio_fullMCTS); cfg.universes = int(a.universes)
        cfg.forced_playouts = int(bool(a.forced_playouts)); cfg.no_mem_optim = int(bool(a.no_mem_optim))
        cfg.dirichlet_noise = int(bool(dirichlet_noise)); cfg.node_cap = node_cap; cfg.edge_cap = edge_cap
        cfg.cpuct = float(a.cpuct); cfg.fpu = float(a.fpu); cfg.dirichletAlpha = float(a.dirichletAlpha)
        cfg.prob_fullMCTS = float(a.prob_fullMCTS)
        t = list(a.temperature) + [1.0] * (3 - len(a.temperature))
        for i in range(3):
            cfg.temperature[i] = float(t[i])
        cfg.tempThreshold = float(a.tempThreshold); cfg.seed = int(seed)
        self.cfg = cfg
        self.h = C.c_void_p()
        _lib.check(self._L.azg_engine_create(C.byref(cfg), nnet.net.h, C.byref(self.h)))
        self.info = game.info

    def reset(self, game=-1):
        _lib.check(self._L.azg_engine_reset(self.h, game))

    def search(self, roots, full_search=None, noise=None, out=None, stream=None):
        """getActionProb for games [0,n). Host (numpy) or device (torch CUDA) buffers.
        Returns (counts int32[n,A], raw_counts int32[n,A], q float32[n,np])."""
        A, S, NPL = self.info.action_size, self.info.state_bytes, self.game.num_players
        if isinstance(roots, np.ndarray):
            roots = np.ascontiguousarray(roots, dtype=np.int8).reshape(-1, S)
            n = len(roots)
            if out is not None:
                counts, raw, q = out
            else:
                counts = np.empty((n, A), np.int32); raw = np.empty((n, A), np.int32); q = np.empty((n, NPL), np.float32)
            fs = None if full_search is None else np.ascontiguousarray(np.asarray(full_search).astype(np.uint8))
            nz = None
            if noise is not None:
                nz = np.zeros((n, A), np.float64)
                for i, x in enumerate(noise):
                    nz[i, :len(x)] = x
        else:                                   # torch CUDA tensors, outputs supplied by the caller
            n = roots.shape[0]
            counts, raw, q = out
            fs, nz = full_search, noise
        _lib.check(self._L.azg_engine_search(self.h, n, _lib.ptr(roots), _lib.ptr(fs), _lib.ptr(nz), _lib.ptr(counts), _lib.ptr(raw),
                                             _lib.ptr(q), stream))
        return counts, raw, q

    def selfplay(self, min_episodes=0, max_moves=0, stream=None):
        _lib.check(self._L.azg_engine_selfplay(self.h, int(min_episodes), int(max_moves), stream))

    def profile(self, enable):
        _lib.check(self._L.azg_engine_profile(self.h, int(bool(enable))))

    def kernel_times(self):
        out = np.zeros(8, np.float64)
        _lib.check(self._L.azg_engine_kernel_times(self.h, _lib.ptr(out)))
        keys = ('select_ms', 'net_ms', 'backup_ms', 'other_ms', 'steps', 'select_launches', 'net_launches', 'backup_launches')
        return dict(zip(keys, out.tolist()))

    def examples(self, cap):
        A, S, NPL = self.info.action_size, self.info.state_bytes, self.game.num_players
        b = np.empty((cap, S), np.int8); pi = np.empty((cap, A), np.float32); z = np.empty((cap, NPL), np.float32)
        va = np.empty((cap, A), np.uint8); q = np.empty((cap, NPL), np.float32); n = C.c_int32(0)
        _lib.check(self._L.azg_engine_examples(self.h, cap, _lib.ptr(b), _lib.ptr(pi), _lib.ptr(z), _lib.ptr(va), _lib.ptr(q), C.byref(n)))
        m = n.value
        return b[:m].reshape((m,) + self.game.getBoardSize()), pi[:m], z[:m], va[:m].astype(np.bool_), q[:m]

    def stats(self):
        out = np.zeros(_lib.AZG_N_STATS, np.int64)
        _lib.check(self._L.azg_engine_stats(self.h, _lib.ptr(out)))
        return dict(zip(_lib.STAT_NAMES, out.tolist()))

    def close(self):
        if self.h:
            self._L.azg_engine_destroy(self.h); self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MCTS:
    """One search tree with the reference's call signature (MCTS.py:24,49)."""

    _instances = []

    def __init__(self, game, nnet, args, dirichlet_noise=False, batch_info=None, seed=0, node_cap=0):
        self.game = game; self.nnet = nnet; self.args = with_defaults(args); self.dirichlet_noise = dirichlet_noise
        self.rng = np.random.default_rng()
        self.engine = Engine(game, nnet, self.args, 1, dirichlet_noise=dirichlet_noise, seed=seed, node_cap=node_cap)
        self.step = 0
        MCTS._instances.append(self)

    def getActionProb(self, canonicalBoard, temp=1, force_full_search=False, noise=None):
        """Returns (probs list[A], q list[np], is_full_search) like MCTS.py:49-103."""
        a = self.args
        is_full = bool(force_full_search or (self.rng.random() < a.prob_fullMCTS))
        nz = None if noise is None or len(noise) == 0 else [noise]
        counts, raw, q = self.engine.search(np.asarray(canonicalBoard)[None], full_search=[is_full], noise=nz)
        counts = counts[0].astype(np.float64)
        self.last_raw_counts = raw[0]
        self.step = (a.numMCTSSims if is_full else a.numMCTSSims // a.ratio_fullMCTS) - 1
        if temp <= 0.02:                                          # MCTS.py:93-98
            best = np.flatnonzero(counts == counts.max())
            probs = [0] * len(counts); probs[int(np.random.choice(best))] = 1
            return probs, [float(x) for x in q[0]], is_full
        c = counts ** (1.0 / temp)
        return list(c / c.sum()), [np.float32(x) for x in q[0]], is_full

    @staticmethod
    def reset_all_search_trees():                                 # MCTS.py:199-203
        for m in MCTS._instances:
            m.engine.reset()
