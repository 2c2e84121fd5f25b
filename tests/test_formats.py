"""On-disk formats: checkpoint.examples written by the reference's own Coach.saveTrainExamples are read, re-written byte
for byte, and converted to arrays; .pt checkpoints of the reference are read without importing its model classes; checkpoints
written here are read back by the reference's own load_checkpoint (that last part needs /root/reference: build container only)."""
import os
import pickle
import sys
import types

import numpy as np
import pytest

from azg_b200 import formats as F
from azg_b200.nnet import V80_TENSOR_ORDER, random_v80_state_dict, v80_blob
from conftest import GOLDEN, load_selfplay_golden

REF = '/root/reference'


@pytest.mark.parametrize('tag,no_compression', [('plain', True), ('zlib', False)])
def test_examples_file_roundtrip(tmp_path, tag, no_compression):
    src = os.path.join(GOLDEN, f'santorini_{tag}.examples')
    hist = F.load_train_examples(src, no_compression=no_compression)
    assert len(hist) == 2 and len(hist[0]) == 10 and len(hist[1]) == 14
    first = hist[0][0]
    assert isinstance(first, tuple) == no_compression
    out = F.save_train_examples(hist, str(tmp_path))
    a = pickle.load(open(out, 'rb')); b = pickle.load(open(src, 'rb'))        # what the reference's loadTrainExamples does (Coach.py:240-241)
    assert type(a) is type(b) is list and [type(h).__name__ for h in a] == ['deque', 'deque'] and [h.maxlen for h in a] == [h.maxlen for h in b]
    for ha, hb in zip(a, b):
        for ea, eb in zip(ha, hb):
            if no_compression:
                assert all(np.array_equal(np.asarray(x), np.asarray(y)) and np.asarray(x).dtype == np.asarray(y).dtype for x, y in zip(ea, eb))
            else:
                assert ea == eb                                               # identical zlib blobs
    if not no_compression:
        assert open(out, 'rb').read() == open(src, 'rb').read()               # the compressed form is byte-identical to the reference's file


def test_examples_harmonise_and_arrays():
    cfg, games = load_selfplay_golden('santorini'); gd = games[3]
    hist = F.load_train_examples(os.path.join(GOLDEN, 'santorini_zlib.examples'), no_compression=True)     # Coach.py:248-251
    assert isinstance(hist[0][0], tuple)
    b, pi, z, va, q = F.examples_to_arrays(list(hist[0]) + list(hist[1]))
    assert (b == gd['ex_board'][:24]).all() and (pi == gd['ex_pi'][:24]).all() and (z == gd['ex_z'][:24]).all()
    assert (va == gd['ex_valids'][:24]).all() and (q == gd['ex_q'][:24]).all()
    plain = F.load_train_examples(os.path.join(GOLDEN, 'santorini_plain.examples'), no_compression=False)  # Coach.py:243-246
    assert isinstance(plain[0][0], bytes) and plain[0][0] == F.compress_example(hist[0][0])
    dq = F.arrays_to_examples(b, pi, z, va, q)
    assert all(np.array_equal(np.asarray(x), np.asarray(y)) and np.asarray(x).dtype == np.asarray(y).dtype for x, y in zip(dq[3], hist[0][3]))
    assert isinstance(dq[3][4], list) and type(dq[3][4][0]) is np.float32          # q stays a list of float32 (MCTS.py:71-72)
    trimmed = F.load_train_examples(os.path.join(GOLDEN, 'santorini_plain.examples'), num_iters_history=1, maxlen_of_queue=5)
    assert len(trimmed) == 1 and len(trimmed[0]) == 5


def test_checkpoint_roundtrip_and_safe_loader(tmp_path):
    sd = random_v80_state_dict(3)
    F.save_checkpoint_file(sd, 80, str(tmp_path), 'temp.pt', additional_keys={'cpuct': 0.8, 'numMCTSSims': 800})
    ck = F.load_checkpoint_file(str(tmp_path / 'temp.pt'))
    assert ck['nn_version'] == 80 and ck['cpuct'] == 0.8
    assert (v80_blob(ck['state_dict']) == v80_blob(sd)).all()
    assert 'first_layer.norm.num_batches_tracked' in ck['state_dict']
    # a pickle that names a dangerous builtin is refused, an unknown class becomes an inert stand-in
    import io
    class Evil:
        def __reduce__(self):
            return (eval, ('1+1',))
    with pytest.raises(pickle.UnpicklingError):
        F._SafeUnpickler(io.BytesIO(pickle.dumps(Evil()))).load()
    mod = types.ModuleType('not_importable_anywhere'); sys.modules['not_importable_anywhere'] = mod
    class Thing:
        pass
    Thing.__module__ = 'not_importable_anywhere'; Thing.__qualname__ = 'Thing'; mod.Thing = Thing
    t = Thing(); t.version = 7
    blob = pickle.dumps(t); del sys.modules['not_importable_anywhere']
    got = F._SafeUnpickler(io.BytesIO(blob)).load()
    assert got.version == 7 and isinstance(got, F._Inert)


@pytest.mark.skipif(not os.path.isdir(REF), reason='needs the reference tree (build container only)')
def test_reference_checkpoints_load_without_reference_on_path():
    assert not any(p.rstrip('/') == REF for p in sys.path)
    for rel, order, ver in (('splendor/pretrained_2players.pt', V80_TENSOR_ORDER, 80),):
        ck = F.load_checkpoint_file(os.path.join(REF, rel))
        assert ck['nn_version'] == ver and all(k in ck['state_dict'] for k in order)
        assert ck['cpuct'] == 0.8 and ck['universes'] == 3                              # the MCTS args stored with the shipped net


@pytest.mark.skipif(not os.path.isdir(REF), reason='needs the reference tree (build container only)')
def test_checkpoint_written_here_loads_in_the_reference(tmp_path):
    """GenericNNetWrapper.load_checkpoint of the UNMODIFIED reference reads a checkpoint written by save_checkpoint_file."""
    import subprocess
    sd = random_v80_state_dict(5)
    F.save_checkpoint_file(sd, 80, str(tmp_path), 'temp.pt', additional_keys={'cpuct': 1.25})
    np.savez(tmp_path / 'want.npz', **sd)
    code = f'''
import sys, os
sys.path[:0] = [{os.path.join(os.path.dirname(GOLDEN), '..', 'oracle', 'ref_shim')!r}, {REF!r}]
os.environ.setdefault('NUMBA_CACHE_DIR', '/tmp/numba_cache')
import numpy as np
from splendor.SplendorGame import SplendorGame
from splendor.NNet import NNetWrapper
w = NNetWrapper(SplendorGame(), dict(nn_version=80, dropout=0., lr=3e-4, learn_rate=3e-4, epochs=2, batch_size=32, no_compression=True, q_weight=0.5))
ck = w.load_checkpoint({str(tmp_path)!r}, 'temp.pt')
assert ck is not None and not w.requestKnowledgeTransfer and ck['cpuct'] == 1.25
want = np.load({str(tmp_path / 'want.npz')!r})
got = w.nnet.state_dict()
assert all((got[k].numpy() == want[k]).all() for k in want.files), 'weights differ'
print('REF_LOAD_OK')
'''
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600)
    assert 'REF_LOAD_OK' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
