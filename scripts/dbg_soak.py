"""Soak (GAME=splendor|santorini|abalone|azul, N, SIMS, EPISODES): complete self-play games at the HEADLINE search budget (800 sims per move, 16 384 games in flight, universes 3, every
move a full search): until n_games episodes have finished, ring drained device-to-device. Prints sims/s over whole games and the
counters that must stay 0 (examples_dropped, arena_overflows, gc_sweeps)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import azg_b200
from azg_b200.mcts import Engine
from azg_b200.utils import dotdict
n = int(os.environ.get('N', 16384)); sims = int(os.environ.get('SIMS', 800)); eps = int(os.environ.get('EPISODES', n))
from azg_b200.game_switcher import import_game, DEFAULT_NN_VERSION
gname = os.environ.get('GAME', 'splendor')
npl = int(os.environ.get('NUM_PLAYERS', 0)) or None
Game, NNet, _ = import_game(gname, npl); game = Game(); net = NNet(game, {'nn_version': DEFAULT_NN_VERSION[gname]})
if os.environ.get('WEIGHTS') == 'shipped':                       # the reference's shipped checkpoint as recorded in the golden vectors
    import numpy as np
    tag = {'splendor': {None: 'splendor_v80_shipped', 2: 'splendor_v80_shipped', 3: 'splendor3p_v80_shipped', 4: 'splendor4p_v80_shipped'}[npl], 'santorini': 'santorini_v89_shipped', 'abalone': 'abalone_v21_shipped', 'azul': 'azul_v84_shipped'}[gname]
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', tag + '.npz'))
    net.load_state_dict({k[4:]: z[k] for k in z.files if k.startswith('sd__')})
    gname += ' (shipped weights)'
a = dotdict(numMCTSSims=sims, cpuct=1.25, fpu=0.0, universes={'splendor': 3, 'azul': 2}.get(gname, 1), dirichletAlpha=-1.0, temperature=[1.0, 0.1, 1.1], tempThreshold=10, prob_fullMCTS=float(os.environ.get('PROB', 1.0)),
            ratio_fullMCTS=5, forced_playouts=False, no_mem_optim=False)
dev = torch.device('cuda', 0)
eng = Engine(game, net, a, n_games=n, dirichlet_noise=True, seed=7, node_cap=int(os.environ.get('NODE_CAP', 0)))
s0 = eng.stats(); torch.cuda.synchronize(); t0 = time.perf_counter(); n_ex = 0
while True:
    left = eps - (eng.stats()['episodes_finished'] - s0['episodes_finished'])
    if left <= 0: break
    eng.selfplay(min_episodes=left)
    n_ex += len(eng.examples_device(dev)[0])
torch.cuda.synchronize(); wall = time.perf_counter() - t0
s1 = eng.stats(); d = {k: s1[k] - s0[k] for k in ('sims', 'moves_played', 'episodes_finished', 'examples_recorded', 'terminal_hits', 'gc_runs', 'gc_sweeps', 'arena_overflows', 'node_visits')}
print(gname, ('%d players' % npl) if npl else '', 'prob_fullMCTS', a.prob_fullMCTS, 'whole games at %d sims: %.1f s, %.2f M sims/s, %d episodes, %d examples (%d drained), mean depth %.2f, moves per game %.1f' % (
    sims, wall, d['sims'] / wall / 1e6, d['episodes_finished'], d['examples_recorded'], n_ex, d['node_visits'] / max(d['sims'], 1), d['moves_played'] / max(d['episodes_finished'], 1)))
print({k: d[k] for k in ('terminal_hits', 'gc_runs', 'gc_sweeps', 'arena_overflows')}, 'gc_trims', s1['gc_trims'] - s0['gc_trims'], 'examples_dropped', s1['examples_dropped'], 'max_nodes', s1.get('max_nodes'), 'node_cap', s1.get('node_cap'))
