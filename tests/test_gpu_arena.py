"""Arena (Arena.py:35-140) / accept gate (Coach.py:194-215) on the engine: every game of an MCTS-vs-MCTS contest in flight at once,
checked game for game against the same contest played sequentially with the CPU oracle's MCTS and rules (separate tree per player,
reused along a game; temp_for_game argmax; 1-2-2-1 seats)."""
import numpy as np
import pytest

import azg_b200
from azg_b200.arena import Arena, EngineArena, accept_new_net, one_vs_two, temp_for_game
from azg_b200.nnet import HashNetWrapper
from azg_b200.utils import with_defaults
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def oracle_contest(game_fns, cfg, args, init_boards, first_index=0):
    ended, canonical, next_state = game_fns
    out = []
    for j, b0 in enumerate(init_boards):
        ovt = one_vs_two(first_index + j)
        trees = [O.MCTS(cfg, None, dirichlet_noise=False, seed=1), O.MCTS(cfg, None, dirichlet_noise=False, seed=2)]
        board, player, it = np.array(b0, copy=True), 0, 0
        while not ended(board, player).any():
            it += 1
            canon = canonical(board, player)
            who = 0 if ((player == 0) == ovt) else 1
            probs, q, full, raw = trees[who].getActionProb(canon, temp=temp_for_game(args, it), force_full_search=True)
            board, player = next_state(board, player, int(np.argmax(probs)))
        out.append((float(ended(board, player)[0]), it))
    return out


def test_engine_arena_matches_sequential_oracle_contest_santorini():
    game = azg_b200.SantoriniGame()
    args = with_defaults(dict(numMCTSSims=48, cpuct=1.25, fpu=0.1, universes=1, forced_playouts=True, tempThreshold=40, arenaCompare=10))
    n = 10
    inits = game.init_batch(np.arange(1, n + 1, dtype=np.uint64) * 104729)
    cfg = O.make_cfg(numMCTSSims=48, cpuct=1.25, fpu=0.1, universes=1, forced_playouts=True, net_kind=0, game=O.GAME_SANTORINI)
    want = oracle_contest((O.sant_game_ended, O.sant_canonical, O.sant_next_state), cfg, args, inits)
    net = HashNetWrapper(game)
    ar = EngineArena(game, net, net, args, n_parallel=n, node_cap=4096)
    res = ar.play_batch(0, n, init_boards=inits)
    st = [e.stats() for e in ar.eng]; ar.close()
    assert [float(r) for r in res] == [w[0] for w in want]
    assert sum(s['arena_overflows'] for s in st) == 0
    # each engine searched only the plies of its own player: together exactly one search per ply of every game
    assert sum(s['sims'] for s in st) == 48 * sum(w[1] for w in want)
    # playGames accounting (Arena.py:126-131) on the same results
    one = sum(1 for j, (r, _) in enumerate(want) if r == (1. if one_vs_two(j) else -1.))
    two = sum(1 for j, (r, _) in enumerate(want) if r == (-1. if one_vs_two(j) else 1.))
    assert one + two <= n and accept_new_net(6, 4, 0.55) and not accept_new_net(5, 5, 0.55) and not accept_new_net(0, 0, 0.55)


def test_engine_arena_playgames_and_reference_style_arena_splendor():
    """Chance game, two batches (n_parallel < num), result accounting sums to num; the reference-style Arena with the MCTS facade as
    players plays a game through the scalar Game calls."""
    from azg_b200.mcts import MCTS
    game = azg_b200.SplendorGame(); net = HashNetWrapper(game)
    args = with_defaults(dict(numMCTSSims=16, universes=2, tempThreshold=10))
    ar = EngineArena(game, net, net, args, n_parallel=4, seed=3, node_cap=1024)
    one, two, draws = ar.playGames(6)
    ar.close()
    assert one + two + draws == 6
    m1 = MCTS(game, net, dict(args, prob_fullMCTS=1.0), node_cap=1024); m2 = MCTS(game, net, dict(args, prob_fullMCTS=1.0), node_cap=1024)
    a = Arena(lambda x, n: int(np.argmax(m1.getActionProb(x, temp=temp_for_game(args, n), force_full_search=True)[0])),
              lambda x, n: int(np.argmax(m2.getActionProb(x, temp=temp_for_game(args, n), force_full_search=True)[0])), game)
    r = a.playGame()
    assert r in (1.0, -1.0) or abs(r) < 0.02
    m1.engine.close(); m2.engine.close()
