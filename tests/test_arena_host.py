"""Host logic of Arena (Arena.py:35-140) on a toy game, no GPU: seat rule, 1-2-2-1 alternation, result accounting, accept gate."""
import numpy as np

from azg_b200.arena import Arena, accept_new_net, one_vs_two, temp_for_game
from azg_b200.utils import with_defaults


class CountDown:
    """Players alternately subtract 1 or 2 from a counter; whoever reaches 0 wins. Duck-typed Game.py surface used by Arena."""
    num_players = 2

    def getNumberOfPlayers(self): return 2
    def getInitBoard(self): return np.array([7, 0], np.int8)
    def getGameEnded(self, board, next_player):
        if board[0] > 0:
            return np.zeros(2, np.float32)
        w = int(board[1])                                  # the seat that moved last won
        return np.array([1., -1.] if w == 0 else [-1., 1.], np.float32)
    def getCanonicalForm(self, board, player): return board
    def getValidMoves(self, board, player): return np.array([0, 1, 1 if board[0] >= 2 else 0])
    def getNextState(self, board, player, action, random_seed=0):
        return np.array([board[0] - action, player], np.int8), 1 - player
    def getScore(self, board, p): return 0


def test_arena_seats_and_accounting(monkeypatch):
    from azg_b200 import arena as A
    monkeypatch.setattr(A.MCTS, 'reset_all_search_trees', staticmethod(lambda: None))
    g = CountDown()
    perfect = lambda b, it: 1 if b[0] % 3 == 1 else (2 if b[0] % 3 == 2 else 1)     # wins from 7 when moving first
    ones = lambda b, it: 1
    a = Arena(perfect, ones, g)
    assert a.playGame() == 1.0                                   # player1 in seat 0 wins
    assert a.playGame(other_way=True) == -1.0                    # seat 0 is now `ones`: the perfect player (seat 1) still wins
    one, two, draws = a.playGames(8)
    assert (one, two, draws) == (8, 0, 0)                        # 1 2 2 1 1 2 2 1: player1 wins from either seat, credited to player1 both ways
    assert Arena(ones, perfect, g).playGames(4) == (0, 4, 0)
    assert [one_vs_two(i) for i in range(8)] == [True, False, False, True, True, False, False, True]


def test_gate_and_temperature():
    assert accept_new_net(11, 9, 0.55) and not accept_new_net(10, 10, 0.55) and not accept_new_net(0, 0, 0.5)
    args = with_defaults(dict(tempThreshold=10))
    assert temp_for_game(args, 0) == 0.5 and abs(temp_for_game(args, 10) - 0.25) < 1e-12 and temp_for_game(args, 47) < 0.02 < temp_for_game(args, 46)
