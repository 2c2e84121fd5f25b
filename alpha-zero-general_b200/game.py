"""`Game` facade over the batched CUDA game-step kernels (reference: Game.py:1-162, splendor/SplendorGame.py:11-71).

Scalar calls are n=1 launches of the batched kernels (used by parity tests and by callers such as the reference's
Arena); the engine itself never goes through this class.
"""
import numpy as np

from . import lib as _lib

NUMBER_PLAYERS = 2


class CudaGame:
    """Same 15-method surface as the reference's Game (Game.py). Batched variants carry a `_batch` suffix."""

    game_id = _lib.AZG_GAME_SPLENDOR
    max_score_diff = 15

    def __init__(self, num_players=NUMBER_PLAYERS):
        self.num_players = num_players
        self._L = _lib.load()
        self.info = _lib.game_info(self.game_id, num_players)
        self._seed_ctr = np.random.SeedSequence().entropy & 0xFFFFFFFFFFFF
        self._board = None

    # ---- sizes
    def getBoardSize(self):
        if self.info.state_depth > 1:
            return (self.info.state_rows, self.info.state_cols, self.info.state_depth)
        return (self.info.state_rows, self.info.state_cols)

    def getActionSize(self):
        return self.info.action_size

    def getNumberOfPlayers(self):
        return self.num_players

    def getMaxScoreDiff(self):
        return self.max_score_diff

    # ---- batched primitives
    def _boards(self, boards):
        b = np.ascontiguousarray(boards, dtype=np.int8).reshape(-1, self.info.state_bytes)
        return b

    def init_batch(self, seeds):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        out = np.empty((len(seeds), self.info.state_bytes), np.int8)
        _lib.check(self._L.azg_game_init(self.game_id, self.num_players, len(seeds), _lib.ptr(seeds), _lib.ptr(out), None))
        return out.reshape((-1,) + self.getBoardSize())

    def valid_batch(self, boards, players=None):
        b = self._boards(boards); n = len(b)
        pl = None if players is None else np.ascontiguousarray(players, dtype=np.int32)
        out = np.empty((n, self.info.action_size), np.uint8)
        _lib.check(self._L.azg_game_valid(self.game_id, self.num_players, n, _lib.ptr(b), _lib.ptr(pl), _lib.ptr(out), None))
        return out.astype(np.bool_)

    def next_batch(self, boards, players, actions, seeds, rng_keys=None):
        b = self._boards(boards); n = len(b)
        pl = np.ascontiguousarray(players, dtype=np.int32); ac = np.ascontiguousarray(actions, dtype=np.int32)
        sd = np.ascontiguousarray(seeds, dtype=np.int64)
        rk = None if rng_keys is None else np.ascontiguousarray(rng_keys, dtype=np.uint64)
        out = np.empty_like(b); onp = np.empty(n, np.int32)
        _lib.check(self._L.azg_game_next(self.game_id, self.num_players, n, _lib.ptr(b), _lib.ptr(pl), _lib.ptr(ac), _lib.ptr(sd),
                                         _lib.ptr(rk), _lib.ptr(out), _lib.ptr(onp), None))
        return out.reshape((-1,) + self.getBoardSize()), onp

    def ended_batch(self, boards, next_players=None):
        b = self._boards(boards); n = len(b)
        npl = None if next_players is None else np.ascontiguousarray(next_players, dtype=np.int32)
        out = np.empty((n, self.num_players), np.float32)
        _lib.check(self._L.azg_game_ended(self.game_id, self.num_players, n, _lib.ptr(b), _lib.ptr(npl), _lib.ptr(out), None))
        return out

    def canonical_batch(self, boards, players):
        b = self._boards(boards); n = len(b)
        pl = np.ascontiguousarray(players, dtype=np.int32); out = np.empty_like(b)
        _lib.check(self._L.azg_game_canonical(self.game_id, self.num_players, n, _lib.ptr(b), _lib.ptr(pl), _lib.ptr(out), None))
        return out.reshape((-1,) + self.getBoardSize())

    def round_score_batch(self, boards):
        b = self._boards(boards); n = len(b)
        rounds = np.empty(n, np.int32); scores = np.empty((n, self.num_players), np.int32)
        _lib.check(self._L.azg_game_round_score(self.game_id, self.num_players, n, _lib.ptr(b), _lib.ptr(rounds), _lib.ptr(scores), None))
        return rounds, scores

    def symmetries_batch(self, boards, pis, valids):
        b = self._boards(boards); n = len(b); K = self.info.max_symmetries; A = self.info.action_size
        pi = np.ascontiguousarray(pis, dtype=np.float32).reshape(n, A)
        va = np.ascontiguousarray(np.asarray(valids).astype(np.uint8)).reshape(n, A)
        ob = np.empty((n, K, self.info.state_bytes), np.int8); op = np.empty((n, K, A), np.float32); ov = np.empty((n, K, A), np.uint8)
        ok = np.empty(n, np.int32)
        _lib.check(self._L.azg_game_symmetries(self.game_id, self.num_players, n, _lib.ptr(b), _lib.ptr(pi), _lib.ptr(va), _lib.ptr(ob),
                                               _lib.ptr(op), _lib.ptr(ov), _lib.ptr(ok), None))
        return ob.reshape((n, K) + self.getBoardSize()), op, ov.astype(np.bool_), ok

    # ---- the reference's scalar surface (Game.py)
    def getInitBoard(self):
        self._seed_ctr += 1
        self._board = self.init_batch([self._seed_ctr])[0]
        return self._board

    def getNextState(self, board, player, action, random_seed=0):
        self._seed_ctr += 1
        nb, npl = self.next_batch(board[None], [player], [action], [random_seed], [self._seed_ctr])
        return nb[0], int(npl[0])

    def getValidMoves(self, board, player):
        return self.valid_batch(board[None], [player])[0]

    def getGameEnded(self, board, next_player):
        return self.ended_batch(board[None], [next_player])[0]

    def getScore(self, board, player):
        return int(self.round_score_batch(board[None])[1][0, player])

    def getRound(self, board):
        return int(self.round_score_batch(board[None])[0][0])

    def getCanonicalForm(self, board, player):
        if player == 0:
            return board
        return self.canonical_batch(board[None], [player])[0]

    def getSymmetries(self, board, pi, valid_actions):
        ob, op, ov, ok = self.symmetries_batch(board[None], np.array(pi, dtype=np.float32)[None], np.asarray(valid_actions)[None])
        return [(ob[0, k], op[0, k], ov[0, k]) for k in range(int(ok[0]))]

    def stringRepresentation(self, board):
        return board.tobytes()

    def moveToString(self, move, current_player):
        return f'action {move}'

    def printBoard(self, numpy_board):
        print(np.asarray(numpy_board))


def splendor_move_to_str(move):
    """Plain-text action names for the 81 Splendor actions (action map: SplendorLogicNumba.py:53-84)."""
    if move < 12:
        return f'buy visible card tier {move // 4} slot {move % 4}'
    if move < 24:
        return f'reserve visible card tier {(move - 12) // 4} slot {(move - 12) % 4}'
    if move < 27:
        return f'reserve blind from deck tier {move - 24}'
    if move < 30:
        return f'buy own reserved card {move - 27}'
    if move < 55:
        return f'take different gems (combination {move - 30})'
    if move < 60:
        return f'take 2 gems of colour {move - 55}'
    if move < 75:
        return f'give back different gems (combination {move - 60})'
    if move < 80:
        return f'give back 2 gems of colour {move - 75}'
    return 'pass'


class SplendorGame(CudaGame):
    """Drop-in for splendor/SplendorGame.py:SplendorGame. The reference fixes the player count with the module constant
    NUMBER_PLAYERS (SplendorGame.py:9); here it is a constructor argument: 2 (default), 3 or 4."""

    def __init__(self, num_players=NUMBER_PLAYERS):
        super().__init__(num_players)

    def moveToString(self, move, current_player):
        return splendor_move_to_str(move)


class SantoriniGame(CudaGame):
    """Drop-in for santorini/SantoriniGame.py:SantoriniGame built without god powers (NB_GODS = 1,
    santorini/SantoriniConstants.py:19): int8[5,5,3] boards, 162 actions = 81*worker + 9*move_direction + build_direction."""

    game_id = _lib.AZG_GAME_SANTORINI
    max_score_diff = 3

    def __init__(self):
        super().__init__(2)

    def moveToString(self, move, current_player):
        dirs = ('NW', 'N', 'NE', 'W', '-', 'E', 'SW', 'S', 'SE')
        return f'worker {move // 81 + 1} moves {dirs[(move % 81) // 9]} and builds {dirs[move % 9]}'


class AbaloneGame(CudaGame):
    """Drop-in for abalone/AbaloneGame.py:AbaloneGame (Belgian daisy start, no dynamic komi: the shipped constants of
    abalone/AbaloneLogicNumba.py:5-6): int8[9,9,4] axial-hex boards, 3402 actions = 378 r + 42 q + plane."""

    game_id = _lib.AZG_GAME_ABALONE
    max_score_diff = 6

    def __init__(self):
        super().__init__(2)

    def moveToString(self, move, current_player):
        plane = move % 42; q = (move // 42) % 9; r = move // 378
        dirs = ('E', 'SE', 'SW', 'W', 'NW', 'NE')
        if plane < 6:
            return f'marble ({r},{q}) moves {dirs[plane]}'
        size, axis = (2, (plane - 6) // 6) if plane < 24 else (3, (plane - 24) // 6)
        return f'{size} marbles from ({r},{q}) along {dirs[axis]} move {dirs[plane % 6]}'


class AzulGame(CudaGame):
    """Drop-in for azul/AzulGame.py:AzulGame (2 players): int8[23,6] boards, 180 actions = 30 source + 6 colour + line
    (source 0 = centre, 1-5 = factories; line 5 = floor), azul/AzulLogicNumba.py:27-48."""

    game_id = _lib.AZG_GAME_AZUL
    max_score_diff = 50

    def __init__(self):
        super().__init__(2)

    def moveToString(self, move, current_player):
        src, colour, line = move // 30, (move % 30) // 6, move % 6
        return f'take colour {colour} from {"the centre" if src == 0 else f"factory {src}"} to {"the floor" if line == 5 else f"line {line + 1}"}'
