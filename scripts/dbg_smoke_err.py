import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import azg_b200
from azg_b200.nnet import NNetWrapper
from oracle import oracle as O
game = azg_b200.SplendorGame()
boards = game.init_batch(np.arange(1, 9, dtype=np.uint64)); valids = game.valid_batch(boards)
kat = np.load('tests/golden/splendor_kat.npz')
for seed in (0, 1, 2):
    for kern in ('tc', 'fp32'):
        if kern == 'fp32': os.environ['AZG_V80_KERNEL'] = 'fp32'
        else: os.environ.pop('AZG_V80_KERNEL', None)
        net = NNetWrapper(game, {'nn_version': 80}, seed=seed)
        blob = O.v80_blob(net.state_dict)
        for name, b, va in (('init', boards, valids), ('kat', kat['canonical'][:200], kat['valids'][:200])):
            pi, v = net.predict_batch(b, va); opi, ov = O.v80_forward(blob, b, va)
            print('seed', seed, kern, name, 'pi err %.3e  v err %.3e  |v| max %.3f' % (np.abs(pi - opi).max(), np.abs(v - ov).max(), np.abs(ov).max()), flush=True)
