#!/usr/bin/env python
"""Golden `checkpoint.examples` files written by the reference's OWN Coach.saveTrainExamples (Coach.py:220-226), compressed and not
(test infrastructure; runs only where /root/reference exists).  python oracle/gen_golden_formats.py [--out tests/golden]"""
import argparse
import os
import sys
import tempfile
from collections import deque

HERE = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault('NUMBA_CACHE_DIR', '/tmp/numba_cache')
sys.path[:0] = [os.path.join(HERE, 'ref_shim'), '/root/reference', HERE, os.path.join(HERE, '..', 'tests')]

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser(); ap.add_argument('--out', default=os.path.join(HERE, '..', 'tests', 'golden')); a = ap.parse_args()
    import Coach as coach_mod
    from conftest import load_selfplay_golden
    import pickle, zlib, shutil
    cfg, games = load_selfplay_golden('santorini')
    gd = games[3]
    ex = [(gd['ex_board'][i], gd['ex_pi'][i], gd['ex_z'][i], gd['ex_valids'][i], [np.float32(x) for x in gd['ex_q'][i]]) for i in range(24)]

    class dotdict(dict):
        __getattr__ = dict.__getitem__
    for tag, comp in (('plain', False), ('zlib', True)):
        c = coach_mod.Coach.__new__(coach_mod.Coach)
        tmp = tempfile.mkdtemp()
        c.args = dotdict(checkpoint=tmp)
        items = [zlib.compress(pickle.dumps(e), level=1) for e in ex] if comp else list(ex)       # Coach.py:84
        c.trainExamplesHistory = [deque(items[:10], maxlen=1000), deque(items[10:], maxlen=1000)]
        c.saveTrainExamples()                                                                       # the reference's writer
        shutil.copy(os.path.join(tmp, 'checkpoint.examples'), os.path.join(a.out, f'santorini_{tag}.examples'))
        print(tag, os.path.getsize(os.path.join(a.out, f'santorini_{tag}.examples')))


if __name__ == '__main__':
    main()
