"""Debug: where does the wall time of a whole-game self-play iteration (40 sims per move) go -- kernels (per-kind event times) vs host."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import azg_b200
from azg_b200.mcts import Engine
from azg_b200.nnet import NNetWrapper
from azg_b200.utils import dotdict
game = azg_b200.SplendorGame(); net = NNetWrapper(game, {'nn_version': 80})
n = int(os.environ.get('N', 16384)); sims = 40
a = dotdict(numMCTSSims=sims, cpuct=1.25, fpu=0.0, universes=3, dirichletAlpha=-1.0, temperature=[1.0, 0.1, 1.1], tempThreshold=10, prob_fullMCTS=1.0,
            ratio_fullMCTS=5, forced_playouts=False, no_mem_optim=False)
dev = torch.device('cuda', 0)
for prof in (0, 1):
    eng = Engine(game, net, a, n_games=n, dirichlet_noise=True, seed=2000, node_cap=8 * sims + 256)
    eng.selfplay(max_moves=2)
    if prof: eng.profile(1)
    s0 = eng.stats(); torch.cuda.synchronize(); t0 = time.perf_counter(); t_sp = t_st = t_ex = 0.0; calls = 0
    while True:
        t = time.perf_counter(); left = n - (eng.stats()['episodes_finished'] - s0['episodes_finished']); t_st += time.perf_counter() - t
        if left <= 0: break
        t = time.perf_counter(); eng.selfplay(min_episodes=left); torch.cuda.synchronize(); t_sp += time.perf_counter() - t; calls += 1
        t = time.perf_counter(); ex = eng.examples_device(dev); torch.cuda.synchronize(); t_ex += time.perf_counter() - t
    wall = time.perf_counter() - t0
    kt = eng.kernel_times() if prof else None
    s1 = eng.stats()
    print('profile' if prof else 'plain  ', 'wall %.3f s | selfplay calls %d: %.3f s | stats %.3f s | examples_device %.3f s | moves %d kernels %d' % (
        wall, calls, t_sp, t_st, t_ex, s1['moves_played'] - s0['moves_played'], s1['kernels_launched'] - s0['kernels_launched']), flush=True)
    if kt: print('   kernel ms:', {k: round(v, 1) for k, v in kt.items() if isinstance(v, float)}, flush=True)
    eng.close()
