"""On-disk formats of the reference, read and written so that both sides can exchange data and weights.

  checkpoint.examples  (Coach.saveTrainExamples / loadTrainExamples, Coach.py:220-262): `pickle.dump` of `trainExamplesHistory`, a list
      (one entry per iteration) of `deque`s of examples; an example is the tuple (board int8[...], pi float32[A], z float32[np],
      valids bool[A], q [float32]*np) or, unless --no-compression, `zlib.compress(pickle.dumps(tuple), level=1)` (Coach.py:84).
  *.pt checkpoints     (GenericNNetWrapper.save_checkpoint / load_checkpoint, GenericNNetWrapper.py:192-277): `torch.save` of a dict
      {'state_dict': OrderedDict[str, Tensor], 'full_model': the pickled nn.Module, **vars(args)}. The reference only reads
      `full_model.version` next to the state_dict on its normal path (load_network, :254-260), so checkpoints written here carry a
      `types.SimpleNamespace(version=V)` there: they load in the reference without this package being importable.
Reading a reference checkpoint does NOT import the reference's model classes: unknown classes found in the pickle are replaced by
inert stand-ins, and only an allow-list of modules (torch, numpy, collections, ...) is resolved at all.
"""
import io
import os
import pickle
import types
import zlib
from collections import OrderedDict, deque

import numpy as np


# ------------------------------------------------------------------ training examples ----------------------
def compress_example(ex):
    """Coach.py:84."""
    return zlib.compress(pickle.dumps(ex), level=1)


def decompress_example(blob):
    """Coach.py:126 / GenericNNetWrapper.pick_examples."""
    return pickle.loads(zlib.decompress(blob))


def harmonise_examples(history, no_compression):
    """Coach.py:243-251: bring loaded examples to the compression mode of the current run (in place)."""
    if not history or not len(history[0]):
        return history
    first = history[0][0]
    if isinstance(first, tuple) and not no_compression:
        for h in history:
            for j in range(len(h)):
                h[j] = compress_example(h[j])
    elif not isinstance(first, tuple) and no_compression:
        for h in history:
            for j in range(len(h)):
                h[j] = decompress_example(h[j])
    return history


def save_train_examples(history, folder, filename='checkpoint.examples'):
    """Coach.saveTrainExamples (Coach.py:220-226)."""
    if not os.path.exists(folder):
        os.makedirs(folder)
    path = os.path.join(folder, filename)
    with open(path, 'wb') as f:
        pickle.dump(history, f)
    return path


def load_train_examples(path, no_compression=True, num_iters_history=None, maxlen_of_queue=None):
    """Coach.loadTrainExamples (Coach.py:228-262): the history list, harmonised and trimmed like the reference does."""
    with open(path, 'rb') as f:
        history = pickle.load(f)
    harmonise_examples(history, no_compression)
    if num_iters_history is not None and len(history) > num_iters_history:
        history = history[-num_iters_history:]
    if maxlen_of_queue is not None:
        for h in history:
            while len(h) > maxlen_of_queue:
                h.pop()
    return history


def examples_to_arrays(examples):
    """List / deque of example tuples (compressed or not) -> (boards int8[n,...], pi f32[n,A], z f32[n,np], valids bool[n,A], q f32[n,np])."""
    ex = [e if isinstance(e, tuple) else decompress_example(e) for e in examples]
    return (np.array([e[0] for e in ex], np.int8), np.array([e[1] for e in ex], np.float32), np.array([e[2] for e in ex], np.float32),
            np.array([e[3] for e in ex], np.bool_), np.array([e[4] for e in ex], np.float32))


def arrays_to_examples(boards, pi, z, valids, q, no_compression=True, maxlen=None):
    """The inverse: a deque of the reference's tuples (what executeEpisodes returns, Coach.py:105-148)."""
    out = deque([], maxlen=maxlen)
    for i in range(len(boards)):
        ex = (np.asarray(boards[i], np.int8), np.asarray(pi[i], np.float32), np.asarray(z[i], np.float32), np.asarray(valids[i], np.bool_),
              [np.float32(x) for x in q[i]])
        out.append(ex if no_compression else compress_example(ex))
    return out


# ------------------------------------------------------------------ .pt checkpoints ------------------------
_ALLOWED_PREFIXES = ('torch', 'numpy', 'collections', '_codecs', 'argparse', 'types', 'copyreg')
_ALLOWED_BUILTINS = {'set', 'frozenset', 'list', 'dict', 'tuple', 'int', 'float', 'bool', 'str', 'bytes', 'bytearray', 'complex', 'slice', 'range', 'object'}


class _Inert:
    """Stand-in for a class the checkpoint pickled but that is not importable here (the reference's nn.Module subclasses):
    it only stores the pickled attributes, so e.g. `full_model.version` stays readable."""
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__['_state'] = state

    def __call__(self, *a, **k):
        return self


class _SafeUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        top = module.split('.')[0]
        if module == 'builtins':
            if name in _ALLOWED_BUILTINS:
                return super().find_class(module, name)
            raise pickle.UnpicklingError(f'refusing builtins.{name} in a checkpoint')
        if top in _ALLOWED_PREFIXES:
            try:
                return super().find_class(module, name)
            except (ImportError, AttributeError):
                pass
        return type(name, (_Inert,), {'__module__': module})


_safe_pickle = types.ModuleType('azg_safe_pickle')
_safe_pickle.Unpickler = _SafeUnpickler
_safe_pickle.Pickler = pickle.Pickler
_safe_pickle.load = lambda f, **kw: _SafeUnpickler(f, **kw).load()
_safe_pickle.loads = lambda b, **kw: _SafeUnpickler(io.BytesIO(b), **kw).load()
_safe_pickle.dump = pickle.dump
_safe_pickle.dumps = pickle.dumps
_safe_pickle.UnpicklingError = pickle.UnpicklingError
_safe_pickle.HIGHEST_PROTOCOL = pickle.HIGHEST_PROTOCOL
_safe_pickle.DEFAULT_PROTOCOL = pickle.DEFAULT_PROTOCOL
_safe_pickle.__name__ = 'pickle'


def load_checkpoint_file(path):
    """torch.load of a reference checkpoint without the reference on sys.path. Returns the dict with 'state_dict' as numpy arrays,
    'nn_version' (from full_model.version if present) and the remaining keys (the args the reference stored) untouched."""
    import torch
    ck = torch.load(path, map_location='cpu', weights_only=False, pickle_module=_safe_pickle)
    out = dict(ck)
    out['state_dict'] = OrderedDict((k, v.detach().cpu().numpy() if hasattr(v, 'detach') else np.asarray(v)) for k, v in ck['state_dict'].items())
    fm = ck.get('full_model')
    out['nn_version'] = getattr(fm, 'version', None)
    return out


def save_checkpoint_file(state_dict, nn_version, folder='checkpoint', filename='checkpoint.pth.tar', additional_keys=None):
    """GenericNNetWrapper.save_checkpoint (GenericNNetWrapper.py:192-205)."""
    import torch
    if not os.path.exists(folder):
        os.mkdir(folder)
    sd = OrderedDict((k, torch.as_tensor(np.asarray(v)).clone() if not hasattr(v, 'detach') else v.detach().cpu().clone()) for k, v in state_dict.items())
    full = OrderedDict()
    if 'lowvalue' not in sd:                                      # the mask constant every net of the reference registers as a buffer (SplendorNNet.py:385)
        full['lowvalue'] = torch.tensor([-1e8], dtype=torch.float32)
    for k, v in sd.items():                                       # BatchNorm buffers the reference's strict load_state_dict expects
        full[k] = v
        nbt = k[:-len('running_var')] + 'num_batches_tracked'
        if k.endswith('running_var') and nbt not in sd:
            full[nbt] = torch.zeros((), dtype=torch.int64)
    data = {'state_dict': full, 'full_model': types.SimpleNamespace(version=int(nn_version))}
    data.update(additional_keys or {})
    path = os.path.join(folder, filename)
    torch.save(data, path)
    return path
