"""Debug: phase breakdown of k_select (library built with -DAZG_SEL_PROF)."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import azg_b200
from azg_b200.mcts import Engine
from azg_b200.nnet import NNetWrapper
game = azg_b200.SplendorGame()
net = NNetWrapper(game, {'nn_version': 80}, seed=0)
args = dict(numMCTSSims=800, cpuct=1.25, fpu=0.0, universes=3, prob_fullMCTS=1.0)
eng = Engine(game, net, args, n_games=16384, node_cap=5120, dirichlet_noise=True)
eng.selfplay(max_moves=3)
L = eng._L; out = (C.c_ulonglong * 16)()
L.azg_debug_selprof(out); a = np.array(list(out), dtype=np.float64)
eng.selfplay(max_moves=1)
L.azg_debug_selprof(out); b = np.array(list(out), dtype=np.float64) - a
n = b[4]
print('walks %d  cycles/walk: total %.0f  root-scan %.0f  materialise %.0f (%.2f calls/walk)  new_leaf %.0f  rest (pointer walk etc) %.0f  depth %.2f' % (
    n, b[0] / n, b[1] / n, b[2] / n, b[5] / n, b[3] / n, (b[0] - b[1] - b[2] - b[3]) / n, b[6] / n))
mx = int(list(out)[7]); c2 = np.array(list(out)[8:], dtype=np.float64)
print('slowest walk so far: %d cycles, depth %d, materialise calls %d; cumulative walks >50k: %d  >100k: %d  >200k: %d  depth>32: %d  depth>64: %d' % (mx >> 20, (mx >> 8) & 0xFFF, mx & 0xFF, c2[2], c2[0], c2[1], c2[3], c2[4]))
print('replay: confirmed levels per walk %.2f of previous path length %.2f (cumulative)' % (c2[5] / (4 * 800 * 16384.0), c2[6] / (4 * 800 * 16384.0)))
print('backup (AZG_SEL_PROF=2): walks %d cycles/warp total %.0f  expansion %.0f  update(phase A) %.0f  refresh(phase B) %.0f' % (b[7], b[0] / max(b[7], 1), b[1] / max(b[7], 1), b[2] / max(b[7], 1), b[3] / max(b[7], 1)))
