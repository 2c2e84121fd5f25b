"""Two fresh engines, same seed, many games in flight: after MOVES self-play moves the boards, players, plies and the root statistics of
every slot must be identical bit for bit (no data race changes a search); a different seed must give different games."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import azg_b200
from azg_b200.mcts import Engine
from azg_b200.game_switcher import import_game, DEFAULT_NN_VERSION
from azg_b200.utils import dotdict
gname = os.environ.get('GAME', 'splendor'); n = int(os.environ.get('N', 16384)); sims = int(os.environ.get('SIMS', 200)); moves = int(os.environ.get('MOVES', 8))
Game, NNet, _ = import_game(gname); game = Game(); net = NNet(game, {'nn_version': DEFAULT_NN_VERSION[gname]})
a = dotdict(numMCTSSims=sims, cpuct=1.25, fpu=0.0, universes={'splendor': 3, 'azul': 2}.get(gname, 1), dirichletAlpha=-1.0, temperature=[1.0, 0.1, 1.1], tempThreshold=10,
            prob_fullMCTS=float(os.environ.get('PROB', 1.0)), ratio_fullMCTS=5, forced_playouts=False, no_mem_optim=False)
def run(seed):
    eng = Engine(game, net, a, n_games=n, dirichlet_noise=True, seed=seed)
    eng.selfplay(max_moves=moves)
    b, pl, ply, act = eng.selfplay_state(); st = eng.stats()
    nd = eng.node(b)                                              # the root statistics of every slot's tree as well (Ns, Nsa, Qsa)
    eng.close()
    return (b, pl, ply, act, nd['Ns'], nd['Nsa'], nd['Qsa']), st
def eq(x, y): return all(np.array_equal(p, q, equal_nan=True) for p, q in zip(x, y))
v1, s1 = run(11); v2, s2 = run(11); v3, _ = run(12)
same = eq(v1, v2) and s1['node_visits'] == s2['node_visits'] and s1['sims'] == s2['sims']
print(gname, 'n', n, 'sims', sims, 'moves', moves, 'prob_full', a.prob_fullMCTS, '| slots', len(v1[0]), '| boards, players, plies, root Ns / Nsa / Qsa identical:', same,
      '| other seed differs:', not eq(v1, v3), '| sims', s1['sims'], s2['sims'], 'node_visits', s1['node_visits'], s2['node_visits'])
sys.exit(0 if same else 1)
