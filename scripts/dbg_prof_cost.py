import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import azg_b200
from azg_b200.mcts import Engine
import bench
game = azg_b200.SplendorGame(); net = azg_b200.NNetWrapper(game, {'nn_version': 80}, seed=0)
a = bench.mcts_args(800, 'splendor')
eng = Engine(game, net, a, n_games=16384, dirichlet_noise=True, seed=1000, node_cap=6*800+320)
eng.selfplay(max_moves=3)
for prof in (False, True, False, True):
    eng.profile(prof)
    s0 = eng.stats(); torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.selfplay(max_moves=3)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    s1 = eng.stats(); kt = eng.kernel_times() if prof else None
    print('profiling', prof, 'sims/s %.2fM' % ((s1['sims'] - s0['sims']) / dt / 1e6), 'us per sim-step %.1f' % (1e6 * dt / 2400), kt and {k: round(v / 2400 * 1e3, 1) for k, v in kt.items() if k.endswith('_ms')})
