"""Arena / accept gate (reference: Arena.py:35-140, Coach.py:194-215,273-276, pit.py:26-64) on the device engine.

Two forms:
  Arena(player1, player2, game, display=None)      the reference's class: any two callables `player(canonical_board, it) -> action`,
        one game at a time, .playGame(verbose, other_way) / .playGames(num) -> (oneWon, twoWon, draws). Same seat rule
        (players = [p1] + [p2]*(np-1), or swapped), same 1-2-2-1 alternation, same result accounting.
  EngineArena(game, nnet1, nnet2, args, n_parallel)  the same contest with BOTH players being MCTS over a net (what Coach.learn and
        pit.py build), all games of the contest in flight at once: one engine per net (separate trees per player like the
        reference's nmcts / pmcts, reused along a game, fresh for every game), every ply = one batched getGameEnded +
        getCanonicalForm + azg_engine_search per engine over the slots whose turn it is (full_search flag 2 idles the others) +
        argmax with the temp_for_game rule + one batched getNextState with a true random chance draw.
"""
import numpy as np

from . import lib as _lib
from .mcts import MCTS, Engine
from .utils import with_defaults


def temp_for_game(args, n):
    """Coach.temp_for_game, Coach.py:273-276."""
    t_begin, t_end, half_life = 0.5, 0.0, abs(args.tempThreshold)
    return t_end + (t_begin - t_end) * (0.5 ** (n / half_life))


def one_vs_two(i):
    """Arena.py:122-125: seat order 1 2 2 1  1 2 2 1 ..."""
    return (i % 4 == 0) or (i % 4 == 3)


class Arena:
    def __init__(self, player1, player2, game, display=None):
        self.player1 = player1; self.player2 = player2; self.game = game; self.display = display

    def playGame(self, initial_state="", verbose=False, other_way=False):
        """One game; returns getGameEnded(board, curPlayer)[0]: 1 = the player in seat 0 won, -1 = lost, else a draw value."""
        g = self.game
        npl = g.getNumberOfPlayers()
        players = ([self.player1] + [self.player2] * (npl - 1)) if not other_way else ([self.player2] + [self.player1] * (npl - 1))
        curPlayer, it = 0, 0
        board = g.getInitBoard()
        if initial_state != "":
            import base64
            import zlib
            data = zlib.decompress(base64.b64decode(initial_state), wbits=-15)
            board = np.frombuffer(data[:-3], dtype=np.int8).reshape(board.shape)
            curPlayer, it = int(data[-3]), int.from_bytes(data[-2:], 'big')
        while not g.getGameEnded(board, curPlayer).any():
            it += 1
            if verbose and self.display:
                self.display(board)
            canonical = g.getCanonicalForm(board, curPlayer)
            action = players[curPlayer](canonical, it)
            valids = g.getValidMoves(canonical, 0)
            assert valids[action] > 0, f'player {curPlayer} chose the illegal action {action}'
            board, curPlayer = g.getNextState(board, curPlayer, action, random_seed=0)
            curPlayer = int(curPlayer)
        if verbose and self.display:
            self.display(board)
        MCTS.reset_all_search_trees()
        return g.getGameEnded(board, curPlayer)[0]

    def playGames(self, num, initial_state="", verbose=False):
        oneWon = twoWon = draws = 0
        for i in range(num):
            ovt = one_vs_two(i) or (initial_state != "")
            r = self.playGame(verbose=verbose, initial_state=initial_state, other_way=not ovt)
            if r == (1. if ovt else -1.):
                oneWon += 1
            elif r == (-1. if ovt else 1.):
                twoWon += 1
            else:
                draws += 1
        return oneWon, twoWon, draws


class EngineArena:
    """All games of an MCTS-vs-MCTS contest at once on the GPU."""

    def __init__(self, game, nnet1, nnet2, args, n_parallel=None, seed=0, node_cap=0, rng=None):
        self.game = game; self.args = a = with_defaults(args)
        self.n = int(n_parallel or a.get('arenaCompare', 30))
        arena_args = dict(a, prob_fullMCTS=1.0)                                  # force_full_search=True, Coach.py:205-206
        # MCTS(game, net, args) in Coach.py:198-201 is built with dirichlet_noise=False
        self.eng = [Engine(game, nnet1, arena_args, self.n, dirichlet_noise=False, seed=seed, node_cap=node_cap),
                    Engine(game, nnet2, arena_args, self.n, dirichlet_noise=False, seed=seed + 1, node_cap=node_cap)]
        self.rng = rng or np.random.default_rng(seed)
        self.key_ctr = (int(seed) << 32) + 1
        self.plies = 0

    def close(self):
        for e in self.eng:
            e.close()

    def _keys(self, n):
        k = np.arange(self.key_ctr, self.key_ctr + n, dtype=np.uint64); self.key_ctr += n
        return k

    def play_batch(self, first_index, count, init_boards=None, chance_seeds=None):
        """Games first_index .. first_index+count-1 (count <= n_parallel) concurrently. Returns the reference's per-game result
        getGameEnded(board, curPlayer)[0] as float32[count]. init_boards / chance_seeds (per ply callables) are test hooks."""
        g = self.game; n = count; A = g.getActionSize()
        idx = np.arange(first_index, first_index + n)
        ovt = np.array([one_vs_two(int(i)) for i in idx])
        boards = (np.asarray(init_boards, np.int8).reshape((n,) + g.getBoardSize()).copy() if init_boards is not None
                  else g.init_batch(self._keys(n)))
        players = np.zeros(n, np.int32); its = np.zeros(n, np.int64)
        result = np.zeros(n, np.float32); alive = np.ones(n, bool)
        for e in self.eng:
            e.reset()                                                            # fresh trees for every game (Arena.py:104)
        pad = self.n - n
        while True:
            ended = g.ended_batch(boards, players)
            done_now = alive & ended.any(axis=1)
            result[done_now] = ended[done_now, 0]
            alive &= ~done_now
            if not alive.any():
                break
            its[alive] += 1
            canon = g.canonical_batch(boards, players)
            # seat 0 belongs to player1 when one_vs_two, else to player2; every other seat to the other one (Arena.py:52-55)
            p1_moves = (players == 0) == ovt
            counts = np.zeros((n, A), np.int64)
            for k, e in enumerate(self.eng):
                mine = alive & (p1_moves if k == 0 else ~p1_moves)
                if not mine.any():
                    continue
                flags = np.where(mine, 1, 2).astype(np.uint8)
                roots = canon
                if pad:
                    flags = np.concatenate([flags, np.full(pad, 2, np.uint8)]); roots = np.concatenate([canon, np.zeros((pad,) + canon.shape[1:], np.int8)])
                c, _, _ = e.search(roots, full_search=flags)
                counts[mine] = c[:n][mine]
            actions = np.zeros(n, np.int32)
            for i in np.flatnonzero(alive):
                # np.argmax(getActionProb(x, temp=temp_for_game(it), force_full_search=True)[0]), Coach.py:205-206 with MCTS.py:93-103:
                # below temp 0.02 a random best action is made one-hot, otherwise counts**(1/temp) keeps the first maximum
                ci = counts[i]
                if temp_for_game(self.args, int(its[i])) <= 0.02:
                    best = np.flatnonzero(ci == ci.max())
                    actions[i] = int(best[0] if len(best) == 1 else self.rng.choice(best))
                else:
                    actions[i] = int(np.argmax(ci))
            seeds = np.zeros(n, np.int64) if chance_seeds is None else np.asarray(chance_seeds(self.plies, n), np.int64)
            nb, npl = g.next_batch(boards, players, actions, seeds, self._keys(n))
            boards = np.where(alive[(...,) + (None,) * (boards.ndim - 1)], nb, boards); players = np.where(alive, npl, players).astype(np.int32)
            self.plies += 1
        return result

    def playGames(self, num):
        """Arena.playGames (Arena.py:107-140): (oneWon, twoWon, draws) over `num` games, 1-2-2-1 seat alternation."""
        oneWon = twoWon = draws = 0
        for first in range(0, num, self.n):
            cnt = min(self.n, num - first)
            res = self.play_batch(first, cnt)
            for j in range(cnt):
                ovt = one_vs_two(first + j)
                if res[j] == (1. if ovt else -1.):
                    oneWon += 1
                elif res[j] == (-1. if ovt else 1.):
                    twoWon += 1
                else:
                    draws += 1
        return oneWon, twoWon, draws


def accept_new_net(nwins, pwins, update_threshold):
    """The gate of Coach.learn (Coach.py:209): reject when no decisive game or the win share is below the threshold."""
    return not (pwins + nwins == 0 or float(nwins) / (pwins + nwins) < update_threshold)
