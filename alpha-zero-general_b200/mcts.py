"""MCTS facade (reference: MCTS.py:19-203) over the batched search engine.

`MCTS(game, nnet, args, dirichlet_noise=False)` + `getActionProb(canonicalBoard, temp, force_full_search)` keep the
reference's signature for one tree; `Engine` exposes the batched form the self-play loop uses.
"""
import ctypes as C
import weakref

import numpy as np

from . import lib as _lib
from .utils import with_defaults


class Engine:
    """n_games concurrent trees on one GPU (azg_engine_*)."""

    def __init__(self, game, nnet, args, n_games, dirichlet_noise=False, seed=0, node_cap=0, edge_cap=0, first_game=0):
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
        cfg.tempThreshold = float(a.tempThreshold); cfg.seed = int(seed); cfg.first_game = int(first_game)
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

    def selfplay_inject(self, init_boards=None, u_full=None, u_move=None, chance_seed=None, noise=None):
        """Replay mode of selfplay() (azg_selfplay_inject): per-slot initial boards [n,S] and per-(slot, ply) random inputs [n,P]
        (noise [n,P,A] optional). Every slot plays ONE game. Call with no arguments to return to the device RNG."""
        if init_boards is None:
            _lib.check(self._L.azg_engine_selfplay_inject(self.h, None)); return
        n, S, A = self.n_games, self.info.state_bytes, self.info.action_size
        ib = np.ascontiguousarray(init_boards, np.int8).reshape(n, S)
        uf = np.ascontiguousarray(u_full, np.float64); um = np.ascontiguousarray(u_move, np.float64); cs = np.ascontiguousarray(chance_seed, np.int64)
        P = uf.shape[1]
        assert uf.shape == (n, P) and um.shape == (n, P) and cs.shape == (n, P)
        nz = None if noise is None else np.ascontiguousarray(noise, np.float64)
        assert nz is None or nz.shape == (n, P, A)
        inj = _lib.SelfplayInject(P, _lib.ptr(ib), _lib.ptr(uf), _lib.ptr(um), _lib.ptr(cs), _lib.ptr(nz))
        _lib.check(self._L.azg_engine_selfplay_inject(self.h, C.byref(inj)))

    def selfplay_state(self):
        """(boards int8[n, *board_shape] absolute frame, players, plies, active) of the self-play slots (Coach.py:55-60 locals)."""
        n, S = self.n_games, self.info.state_bytes
        b = np.empty((n, S), np.int8); pl = np.empty(n, np.int32); ply = np.empty(n, np.int32); act = np.empty(n, np.int32)
        _lib.check(self._L.azg_engine_selfplay_state(self.h, _lib.ptr(b), _lib.ptr(pl), _lib.ptr(ply), _lib.ptr(act)))
        return b.reshape((n,) + self.game.getBoardSize()), pl, ply, act.astype(np.bool_)

    def node(self, boards, slots=None):
        """MCTS.nodes_data[s] (MCTS.py:37-39) for each board, looked up in the tree of slots[i] (default slot i): dict of arrays
        found (0 absent / 1 expanded / 2 terminal), Es, Vs, Ps, Ns, Qsa, Nsa, r, Qs with the reference's dtypes (Nsa int32)."""
        A, S, NPL = self.info.action_size, self.info.state_bytes, self.game.num_players
        b = np.ascontiguousarray(boards, np.int8).reshape(-1, S); n = len(b)
        sl = None if slots is None else np.ascontiguousarray(slots, np.int32)
        o = dict(found=np.empty(n, np.int32), Es=np.empty((n, NPL), np.float32), Vs=np.empty((n, A), np.uint8), Ps=np.empty((n, A), np.float32),
                 Ns=np.empty(n, np.int32), Qsa=np.empty((n, A), np.float64), Nsa=np.empty((n, A), np.int32), r=np.empty(n, np.int32), Qs=np.empty(n, np.float32))
        _lib.check(self._L.azg_engine_node(self.h, n, _lib.ptr(sl), _lib.ptr(b), *[_lib.ptr(o[k]) for k in ('found', 'Es', 'Vs', 'Ps', 'Ns', 'Qsa', 'Nsa', 'r', 'Qs')], None))
        o['Vs'] = o['Vs'].astype(np.bool_)
        return o

    def profile(self, enable):
        """True / 1: time every launch; N > 1: time every N-th lock-step simulation (sampled); False / 0: off."""
        _lib.check(self._L.azg_engine_profile(self.h, int(enable)))

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

    def examples_pending(self):
        n = C.c_int32(0)
        _lib.check(self._L.azg_engine_examples_pending(self.h, C.byref(n)))
        return n.value

    def examples_device(self, device=None):
        """Drains the example ring into torch CUDA tensors (device-to-device: the examples never touch host memory).
        Returns (boards int8[m, *board_shape], pi f32[m,A], z f32[m,np], valids uint8[m,A], q f32[m,np])."""
        import torch
        A, S, NPL = self.info.action_size, self.info.state_bytes, self.game.num_players
        m = self.examples_pending()
        dev = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        b = torch.empty((m, S), dtype=torch.int8, device=dev); pi = torch.empty((m, A), dtype=torch.float32, device=dev)
        z = torch.empty((m, NPL), dtype=torch.float32, device=dev); va = torch.empty((m, A), dtype=torch.uint8, device=dev)
        q = torch.empty((m, NPL), dtype=torch.float32, device=dev); n = C.c_int32(0)
        if m:
            _lib.check(self._L.azg_engine_examples(self.h, m, _lib.ptr(b), _lib.ptr(pi), _lib.ptr(z), _lib.ptr(va), _lib.ptr(q), C.byref(n)))
            assert n.value == m
        return b.view((m,) + self.game.getBoardSize()), pi, z, va, q

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


class _NodesView:
    def __init__(self, mcts):
        self.m = mcts

    def _q(self, key):
        b = np.frombuffer(key, np.int8) if isinstance(key, (bytes, bytearray)) else np.asarray(key, np.int8)
        return {k: v[0] for k, v in self.m.engine.node(b[None]).items()}

    def __contains__(self, key):
        return bool(self._q(key)['found'])

    def __getitem__(self, key):
        o = self._q(key)
        if not o['found']:
            raise KeyError(key)
        if o['found'] == 2:
            return (o['Es'], None, None, 0, None, None, int(o['r']), 0.0)
        return (o['Es'], o['Vs'], o['Ps'], int(o['Ns']), o['Qsa'], o['Nsa'].astype(np.int64), int(o['r']), np.float32(o['Qs']))


class MCTS:
    """One search tree with the reference's call signature (MCTS.py:24,49)."""

    _instances = weakref.WeakSet()        # MCTS.py:199-203 finds the live trees through gc.get_objects(); a WeakSet does the same without pinning them

    def __init__(self, game, nnet, args, dirichlet_noise=False, batch_info=None, seed=0, node_cap=0):
        self.game = game; self.nnet = nnet; self.args = with_defaults(args); self.dirichlet_noise = dirichlet_noise
        self.rng = np.random.default_rng()
        self.engine = Engine(game, nnet, self.args, 1, dirichlet_noise=dirichlet_noise, seed=seed, node_cap=node_cap)
        self.step = 0
        MCTS._instances.add(self)

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

    @property
    def nodes_data(self):
        """Read-only dict-like view of the tree: nodes_data[game.stringRepresentation(board)] -> (Es, Vs, Ps, Ns, Qsa, Nsa, r, Qs)
        as in MCTS.py:37-39 (a terminal node has Vs..Nsa = None like the reference's, MCTS.py:133)."""
        return _NodesView(self)

    @staticmethod
    def reset_all_search_trees():                                 # MCTS.py:199-203
        for m in list(MCTS._instances):
            if m.engine.h:                                        # closed engines have no trees left to reset
                m.engine.reset()
