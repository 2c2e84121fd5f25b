#!/usr/bin/env python
"""Golden fixture for the training step: the reference's OWN GenericNNetWrapper.train (GenericNNetWrapper.py:44-92) on a fixed set of
examples with fixed sample ids (test infrastructure; runs only where /root/reference exists).
    python oracle/gen_golden_train.py [--out tests/golden]
Writes tests/golden/splendor_train_step.npz: initial state_dict, the examples, the sample ids np.random.choice drew, the per-batch
losses and the state_dict after 1 epoch x 3 batches of 16."""
import argparse
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault('NUMBA_CACHE_DIR', '/tmp/numba_cache')
os.environ['OMP_NUM_THREADS'] = '1'
sys.path[:0] = [os.path.join(HERE, 'ref_shim'), '/root/reference', HERE, os.path.join(HERE, '..', 'tests')]

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser(); ap.add_argument('--out', default=os.path.join(HERE, '..', 'tests', 'golden')); a = ap.parse_args()
    torch.set_num_threads(1)
    from conftest import load_selfplay_golden
    from splendor.SplendorGame import SplendorGame
    from splendor.NNet import NNetWrapper
    import GenericNNetWrapper as gw
    cfg, games = load_selfplay_golden('splendor'); gd = games[0]
    n = 48
    sel = np.linspace(0, len(gd['ex_board']) - 1, n).astype(int)
    examples = [(gd['ex_board'][i], gd['ex_pi'][i], gd['ex_z'][i], gd['ex_valids'][i], [np.float32(x) for x in gd['ex_q'][i]]) for i in sel]
    g = SplendorGame()
    torch.manual_seed(0)
    w = NNetWrapper(g, dict(nn_version=80, dropout=0., lr=1e-3, learn_rate=1e-3, epochs=1, batch_size=16, no_compression=True, q_weight=0.5))
    w.device['training'] = 'cpu'
    with torch.no_grad():                                        # non-trivial BatchNorm statistics
        gen = torch.Generator().manual_seed(3)
        for mod in w.nnet.modules():
            if isinstance(mod, torch.nn.BatchNorm1d):
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=gen) * 0.2); mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=gen) + 0.5)
    sd0 = {k: v.detach().clone().numpy() for k, v in w.nnet.state_dict().items()}
    drawn = []
    real_choice = np.random.choice

    def recording_choice(nn_, size=None, replace=True, p=None):
        ids = real_choice(nn_, size=size, replace=replace, p=p); drawn.append(np.array(ids)); return ids
    losses = []
    real_update = gw.AverageMeter.update

    def rec_update(self, val, n=1):
        losses.append(float(val)); return real_update(self, val, n)
    np.random.seed(11)
    np.random.choice = recording_choice; gw.AverageMeter.update = rec_update
    try:
        w.train(examples)                                        # <- the reference's own training loop
    finally:
        np.random.choice = real_choice; gw.AverageMeter.update = real_update
    sd1 = {k: v.detach().clone().numpy() for k, v in w.nnet.state_dict().items()}
    save = {'ids': np.array(drawn), 'losses': np.array(losses).reshape(-1, 2), 'boards': np.array([e[0] for e in examples], np.int8),
            'pi': np.array([e[1] for e in examples], np.float32), 'z': np.array([e[2] for e in examples], np.float32),
            'valids': np.array([e[3] for e in examples], np.bool_), 'q': np.array([e[4] for e in examples], np.float32)}
    for k, v in sd0.items():
        save['sd0__' + k] = v
    for k, v in sd1.items():
        save['sd1__' + k] = v
    np.savez_compressed(os.path.join(a.out, 'splendor_train_step.npz'), **save)
    print('batches', len(drawn), 'losses', save['losses'].tolist())


if __name__ == '__main__':
    main()
