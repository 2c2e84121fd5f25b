"""Inference half of GenericNNetWrapper (GenericNNetWrapper.py:94-157) on the CUDA forward kernels.

`NNetWrapper(game, nn_args)` mirrors splendor/NNet.py:NNetWrapper for nn_version 80. Weights come from a
reference checkpoint / state_dict; torch (if present) is used only to read them.
"""
import ctypes as C

import numpy as np

from . import lib as _lib


def _lin_bn(prefix):
    return [f'{prefix}.linear.weight', f'{prefix}.norm.weight', f'{prefix}.norm.bias',
            f'{prefix}.norm.running_mean', f'{prefix}.norm.running_var']


def _v80_order():
    names = _lin_bn('first_layer')
    for blk in ('trunk.0', 'output_layers_PI.0', 'output_layers_V.0'):
        names += _lin_bn(f'{blk}.expand') + _lin_bn(f'{blk}.depthwise')
        names += [f'{blk}.se.fc1.weight', f'{blk}.se.fc1.bias', f'{blk}.se.fc2.weight', f'{blk}.se.fc2.bias']
        names += _lin_bn(f'{blk}.project')
    names += ['output_layers_PI.2.weight', 'output_layers_PI.2.bias', 'output_layers_PI.4.weight', 'output_layers_PI.4.bias',
              'output_layers_V.2.weight', 'output_layers_V.2.bias', 'output_layers_V.4.weight', 'output_layers_V.4.bias']
    return names


# Order in which azg_net_create expects the SplendorNNet V80 state_dict tensors (splendor/SplendorNNet.py:259-280),
# each flattened row-major and concatenated as float32.
V80_TENSOR_ORDER = _v80_order()


def v80_blob(state_dict):
    parts = []
    for n in V80_TENSOR_ORDER:
        t = state_dict[n]
        if hasattr(t, 'detach'):
            t = t.detach().cpu().numpy()
        parts.append(np.asarray(t, dtype=np.float32).ravel())
    return np.ascontiguousarray(np.concatenate(parts), dtype=np.float32)


def random_v80_state_dict(seed=0, num_players=2):
    """Random-init V80 weights as a freshly constructed reference net really has them, drawn from numpy so no torch is needed.
    The reference's `_init` hook (kaiming_uniform_ weights, zero biases; SplendorNNet.py:385-395) never runs: it walks
    `self.__dict__`, which holds no sub-modules (they live in `_modules`), so every Linear keeps PyTorch's default
    initialisation U(+-1/sqrt(fan_in)) for weights AND biases (the golden random-init state dict confirms it:
    tests/golden/splendor_v80_rand.npz); BatchNorm is the default. Used by bench.py and smoke()."""
    rng = np.random.default_rng(seed)
    nv = 32 + 10 * num_players + num_players * num_players
    E, A = 3 * nv, 81
    Q = max(8, (E // 4 + 4) // 8 * 8); Q += 8 if Q < 0.9 * (E // 4) else 0          # _make_divisible(E // 4, 8): 40 / 56 / 64 for 2 / 3 / 4 players
    shapes = {}
    def lin_bn(prefix, out, inn, ch):
        shapes[f'{prefix}.linear.weight'] = (out, inn)
        for k, val in (('weight', 1.0), ('bias', 0.0), ('running_mean', 0.0), ('running_var', 1.0)):
            shapes[f'{prefix}.norm.{k}'] = ((ch,), val)
    lin_bn('first_layer', nv, nv, nv)
    for blk in ('trunk.0', 'output_layers_PI.0', 'output_layers_V.0'):
        lin_bn(f'{blk}.expand', E, nv, E); lin_bn(f'{blk}.depthwise', 7, 7, E)
        shapes[f'{blk}.se.fc1.weight'] = (Q, E); shapes[f'{blk}.se.fc1.bias'] = ((Q,), 0.0)
        shapes[f'{blk}.se.fc2.weight'] = (E, Q); shapes[f'{blk}.se.fc2.bias'] = ((E,), 0.0)
        lin_bn(f'{blk}.project', nv, E, nv)
    shapes['output_layers_PI.2.weight'] = (A, nv * 7); shapes['output_layers_PI.2.bias'] = ((A,), 0.0)
    shapes['output_layers_PI.4.weight'] = (A, A); shapes['output_layers_PI.4.bias'] = ((A,), 0.0)
    shapes['output_layers_V.2.weight'] = (num_players, nv * 7); shapes['output_layers_V.2.bias'] = ((num_players,), 0.0)
    shapes['output_layers_V.4.weight'] = (num_players, num_players); shapes['output_layers_V.4.bias'] = ((num_players,), 0.0)
    sd = {}
    for name, shp in shapes.items():
        if isinstance(shp[0], tuple):
            sd[name] = np.full(shp[0], shp[1], np.float32)
        else:
            bound = 1.0 / np.sqrt(shp[1])                       # nn.Linear default: kaiming_uniform_(a=sqrt(5)) = U(+-1/sqrt(fan_in))
            sd[name] = rng.uniform(-bound, bound, size=shp).astype(np.float32)
    for name in list(sd):                                       # Linear biases (SE and head linears): U(+-1/sqrt(fan_in)) as well
        if name.endswith('.bias') and '.norm.' not in name:
            fan_in = shapes[name[:-5] + '.weight'][1]
            sd[name] = rng.uniform(-1.0 / np.sqrt(fan_in), 1.0 / np.sqrt(fan_in), size=sd[name].shape).astype(np.float32)
    return sd


def _bn(prefix):
    return [f'{prefix}.weight', f'{prefix}.bias', f'{prefix}.running_mean', f'{prefix}.running_var']


def _v89_order():
    names = ['first_layer.0.weight'] + _bn('first_layer.1')
    for b in range(5):
        names += [f'trunk.{b}.conv1.weight'] + _bn(f'trunk.{b}.bn1') + [f'trunk.{b}.conv2.weight'] + _bn(f'trunk.{b}.bn2')
    names += ['head_PI.conv1x1.weight'] + _bn('head_PI.bn') + ['head_PI.fc.weight', 'head_PI.fc.bias']
    names += ['head_V.conv1x1.weight'] + _bn('head_V.bn') + ['head_V.fc1.weight', 'head_V.fc1.bias', 'head_V.fc2.weight', 'head_V.fc2.bias']
    return names


# Order in which azg_net_create expects the SantoriniNNet V89 state_dict tensors (santorini/SantoriniNNet.py:194-217).
V89_TENSOR_ORDER = _v89_order()


def v89_blob(state_dict):
    parts = []
    for n in V89_TENSOR_ORDER:
        t = state_dict[n]
        if hasattr(t, 'detach'):
            t = t.detach().cpu().numpy()
        parts.append(np.asarray(t, dtype=np.float32).ravel())
    return np.ascontiguousarray(np.concatenate(parts), dtype=np.float32)


def random_v89_state_dict(seed=0):
    """Random-init V89 weights drawn from numpy as a freshly constructed reference net has them: PyTorch-default convolutions and
    Linears (U(+-1/sqrt(fan_in)) for weights and Linear biases; the `_init` hook of SantoriniNNet.py:222-232 never touches a
    layer, see random_v80_state_dict), default BatchNorm. Used by bench.py and smoke()."""
    rng = np.random.default_rng(seed)
    sd = {}

    def conv(name, out, cin, k):
        b = 1.0 / np.sqrt(cin * k * k)
        sd[name] = rng.uniform(-b, b, size=(out, cin, k, k)).astype(np.float32)

    def bn(prefix, ch):
        sd[f'{prefix}.weight'] = np.ones(ch, np.float32); sd[f'{prefix}.bias'] = np.zeros(ch, np.float32)
        sd[f'{prefix}.running_mean'] = np.zeros(ch, np.float32); sd[f'{prefix}.running_var'] = np.ones(ch, np.float32)

    def lin(name, out, inn):
        b = 1.0 / np.sqrt(inn)
        sd[f'{name}.weight'] = rng.uniform(-b, b, size=(out, inn)).astype(np.float32); sd[f'{name}.bias'] = rng.uniform(-b, b, size=out).astype(np.float32)

    conv('first_layer.0.weight', 64, 2, 3); bn('first_layer.1', 64)
    for blk in range(5):
        conv(f'trunk.{blk}.conv1.weight', 64, 64, 3); bn(f'trunk.{blk}.bn1', 64)
        conv(f'trunk.{blk}.conv2.weight', 64, 64, 3); bn(f'trunk.{blk}.bn2', 64)
    conv('head_PI.conv1x1.weight', 2, 64, 1); bn('head_PI.bn', 2); lin('head_PI.fc', 162, 50)
    conv('head_V.conv1x1.weight', 1, 64, 1); bn('head_V.bn', 1); lin('head_V.fc1', 64, 25); lin('head_V.fc2', 2, 64)
    return sd


def _v21_order():
    names = ['first_layer.0.weight'] + _bn('first_layer.1')
    for b in range(4):
        for j in range(3):
            names += [f'trunk.{b}.block.{j}.0.weight'] + _bn(f'trunk.{b}.block.{j}.1')
    names += ['meta_fc.0.weight', 'meta_fc.0.bias', 'head_PI.0.weight'] + _bn('head_PI.1') + ['head_V_conv.0.weight'] + _bn('head_V_conv.1')
    names += ['head_V_fc.0.weight', 'head_V_fc.0.bias', 'head_V_fc.2.weight', 'head_V_fc.2.bias']
    return names


# Order in which azg_net_create expects the AbaloneNNet V21 state_dict tensors (abalone/AbaloneNNet.py:117-156).
V21_TENSOR_ORDER = _v21_order()


def v21_blob(state_dict):
    parts = []
    for n in V21_TENSOR_ORDER:
        t = state_dict[n]
        if hasattr(t, 'detach'):
            t = t.detach().cpu().numpy()
        parts.append(np.asarray(t, dtype=np.float32).ravel())
    return np.ascontiguousarray(np.concatenate(parts), dtype=np.float32)


def random_v21_state_dict(seed=0):
    """Random-init V21 weights from numpy as a freshly constructed reference net has them (PyTorch-default convolutions and Linears:
    U(+-1/sqrt(fan_in)) for weights and Linear biases, see random_v80_state_dict; default BatchNorm). Used by bench.py only."""
    rng = np.random.default_rng(seed)
    sd = {}

    def conv(name, out, cin, k):
        b = 1.0 / np.sqrt(cin * k * k)
        sd[name] = rng.uniform(-b, b, size=(out, cin, k, k)).astype(np.float32)

    def bn(prefix, ch):
        sd[f'{prefix}.weight'] = np.ones(ch, np.float32); sd[f'{prefix}.bias'] = np.zeros(ch, np.float32)
        sd[f'{prefix}.running_mean'] = np.zeros(ch, np.float32); sd[f'{prefix}.running_var'] = np.ones(ch, np.float32)

    def lin(name, out, inn):
        b = 1.0 / np.sqrt(inn)
        sd[f'{name}.weight'] = rng.uniform(-b, b, size=(out, inn)).astype(np.float32); sd[f'{name}.bias'] = rng.uniform(-b, b, size=out).astype(np.float32)

    conv('first_layer.0.weight', 24, 3, 3); bn('first_layer.1', 24)
    for blk in range(4):
        conv(f'trunk.{blk}.block.0.0.weight', 48, 24, 1); bn(f'trunk.{blk}.block.0.1', 48)
        conv(f'trunk.{blk}.block.1.0.weight', 48, 1, 3); bn(f'trunk.{blk}.block.1.1', 48)
        conv(f'trunk.{blk}.block.2.0.weight', 24, 48, 1); bn(f'trunk.{blk}.block.2.1', 24)
    lin('meta_fc.0', 16, 6)
    conv('head_PI.0.weight', 42, 24, 1); bn('head_PI.1', 42)
    conv('head_V_conv.0.weight', 4, 24, 1); bn('head_V_conv.1', 4)
    lin('head_V_fc.0', 64, 340); lin('head_V_fc.2', 2, 64)
    return sd


# AzulNNet V84 (azul/AzulNNet.py:84-111) uses SplendorNNet V80's module names: same tensor order, different shapes.
V84_TENSOR_ORDER = _v80_order()
V84_BLOCKS = (('trunk.0', 23, 115, 23, 32), ('output_layers_PI.0', 23, 115, 46, 32), ('output_layers_V.0', 23, 46, 23, 16))   # name, in, E, out, Q


def v84_blob(state_dict):
    parts = []
    for n in V84_TENSOR_ORDER:
        t = state_dict[n]
        if hasattr(t, 'detach'):
            t = t.detach().cpu().numpy()
        parts.append(np.asarray(t, dtype=np.float32).ravel())
    return np.ascontiguousarray(np.concatenate(parts), dtype=np.float32)


def random_v84_state_dict(seed=0):
    """Random-init V84 weights from numpy as a freshly constructed reference net has them (PyTorch-default Linears, default
    BatchNorm; see random_v80_state_dict)."""
    rng = np.random.default_rng(seed)
    sd = {}

    def lin(name, out, inn, bias):
        b = 1.0 / np.sqrt(inn)
        sd[f'{name}.weight'] = rng.uniform(-b, b, size=(out, inn)).astype(np.float32)
        if bias:
            sd[f'{name}.bias'] = rng.uniform(-b, b, size=out).astype(np.float32)

    def bn(prefix, ch):
        sd[f'{prefix}.weight'] = np.ones(ch, np.float32); sd[f'{prefix}.bias'] = np.zeros(ch, np.float32)
        sd[f'{prefix}.running_mean'] = np.zeros(ch, np.float32); sd[f'{prefix}.running_var'] = np.ones(ch, np.float32)

    lin('first_layer.linear', 23, 23, False); bn('first_layer.norm', 23)
    for name, inn, E, out, Q in V84_BLOCKS:
        lin(f'{name}.expand.linear', E, inn, False); bn(f'{name}.expand.norm', E)
        lin(f'{name}.depthwise.linear', 6, 6, False); bn(f'{name}.depthwise.norm', E)
        lin(f'{name}.se.fc1', Q, E, True); lin(f'{name}.se.fc2', E, Q, True)
        lin(f'{name}.project.linear', out, E, False); bn(f'{name}.project.norm', out)
    lin('output_layers_PI.2', 180, 276, True); lin('output_layers_PI.4', 180, 180, True)
    lin('output_layers_V.2', 2, 138, True); lin('output_layers_V.4', 2, 2, True)
    return sd


class CudaNet:
    """Owns an azg_net handle."""

    def __init__(self, kind, game, weights=None):
        self._L = _lib.load()
        self.kind = kind
        self.game = game
        self.h = C.c_void_p()
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float32)
        _lib.check(self._L.azg_net_create(kind, game.game_id, game.num_players, _lib.ptr(w), 0 if w is None else w.size, C.byref(self.h)))

    def load(self, weights):
        w = np.ascontiguousarray(weights, dtype=np.float32)
        _lib.check(self._L.azg_net_load(self.h, _lib.ptr(w), w.size))

    def forward(self, boards, valids):
        """boards int8 [n,...], valids bool/uint8 [n,A] (numpy, or torch CUDA tensors) -> (pi, v) numpy float32."""
        info = self.game.info
        b = np.ascontiguousarray(boards, dtype=np.int8).reshape(-1, info.state_bytes); n = len(b)
        va = np.ascontiguousarray(np.asarray(valids).astype(np.uint8)).reshape(n, info.action_size)
        pi = np.empty((n, info.action_size), np.float32); v = np.empty((n, self.game.num_players), np.float32)
        _lib.check(self._L.azg_net_forward(self.h, n, _lib.ptr(b), _lib.ptr(va), _lib.ptr(pi), _lib.ptr(v), None))
        return pi, v

    def close(self):
        if self.h:
            self._L.azg_net_destroy(self.h); self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BatchedPredictMixin:
    """predict_client / predict_server (GenericNNetWrapper.py:122-157): the reference's lock-chain protocol between N self-play
    threads and one inference server, unchanged (same shared_memory slots, same lock hand-over), with the batch evaluated by
    `self.predict_batch` (one azg_net_forward over the N boards) instead of an onnxruntime session. `Coach.executeEpisodes_batch`
    threads of the reference can therefore run against this wrapper as they are."""

    def predict_client(self, board, valid_actions, batch_info):
        i_thread, i_result, shared_memory, locks = batch_info
        shared_memory[i_thread] = (np.expand_dims(np.asarray(board), 0), np.expand_dims(np.array(valid_actions).astype(np.bool_), 0))
        locks[i_thread + 1].release()                              # unblock the next thread (= next MCTS or the server) ...
        locks[i_thread].acquire()                                  # ... and wait for our turn
        pi, v = shared_memory[i_result]
        return pi, v

    def predict_server(self, nb_threads, shared_memory, locks):
        locks[0].release()
        while shared_memory[-1] <= 1:
            locks[-1].acquire()                                    # all inputs are in
            boards = np.concatenate([x[0] for x in shared_memory[:nb_threads]])
            valids = np.concatenate([x[1] for x in shared_memory[:nb_threads]])
            pi, v = self.predict_batch(boards, valids)
            for i in range(nb_threads):
                shared_memory[i + nb_threads] = (pi[i], v[i])
            locks[0].release()                                     # unblock the first thread


class NNetWrapper(BatchedPredictMixin):
    """splendor/NNet.py:NNetWrapper (inference surface). nn_args['nn_version'] must be 80."""
    NN_VERSION = 80

    def __init__(self, game, nn_args=None, state_dict=None, seed=0):
        nn_args = dict(nn_args or {'nn_version': 80})
        if nn_args.get('nn_version', 80) != 80:
            raise NotImplementedError('only SplendorNNet version 80 is built (the shipped 2-, 3- and 4-player checkpoints)')
        self.args = nn_args
        self.game = game
        self.board_size = game.getBoardSize(); self.action_size = game.getActionSize(); self.num_players = game.num_players
        self.requestKnowledgeTransfer = False
        self.state_dict = state_dict if state_dict is not None else random_v80_state_dict(seed, game.num_players)
        self.net = CudaNet(_lib.AZG_NET_SPLENDOR_V80, game, v80_blob(self.state_dict))

    # NeuralNet.predict (NeuralNet.py:27, GenericNNetWrapper.py:94-120): single board
    def predict(self, board, valid_actions):
        pi, v = self.net.forward(np.asarray(board)[None], np.asarray(valid_actions)[None])
        return pi[0], v[0]

    # batched form of predict_server (GenericNNetWrapper.py:139-157)
    def predict_batch(self, boards, valid_actions):
        return self.net.forward(boards, valid_actions)

    def load_state_dict(self, state_dict):
        self.state_dict = state_dict
        self.net.load(v80_blob(state_dict))

    def load_checkpoint(self, folder='checkpoint', filename='checkpoint.pth.tar'):
        """GenericNNetWrapper.load_checkpoint (GenericNNetWrapper.py:207-221): reads a reference checkpoint (or one written by
        save_checkpoint) and loads its state_dict into the device net. A missing file prints and returns None like the reference.
        The reference's model classes are NOT imported (formats.load_checkpoint_file stubs them), only 'state_dict' is used."""
        import os
        from .formats import load_checkpoint_file
        path = os.path.join(folder, filename)
        if not os.path.exists(path):
            print('No model in path {}'.format(path))
            return None
        ck = load_checkpoint_file(path)
        if ck.get('nn_version') is not None and self.args.get('nn_version', 0) > 0 and ck['nn_version'] != self.args['nn_version']:
            print('Checkpoint includes NN version', ck['nn_version'], ', but you ask version', self.args['nn_version'], ' so not loading it and initiate knowledge transfer')
            self.requestKnowledgeTransfer = True                   # GenericNNetWrapper.py:254-257
            return None
        self.load_state_dict(ck['state_dict'])
        return ck

    def save_checkpoint(self, folder='checkpoint', filename='checkpoint.pth.tar', additional_keys={}):
        """GenericNNetWrapper.save_checkpoint (GenericNNetWrapper.py:192-205); readable by the reference's load_checkpoint."""
        from .formats import save_checkpoint_file
        return save_checkpoint_file(self.state_dict, self.args.get('nn_version', self.NN_VERSION), folder, filename, additional_keys)

    def train(self, examples, validation_set=None, save_folder=None, every=0):
        """GenericNNetWrapper.train (GenericNNetWrapper.py:44-92): examples = list / deque of the reference's example tuples, or the
        five arrays / CUDA tensors of Engine.examples_device(). Trains with train.Trainer (torch autograd) and loads the new weights
        into the CUDA inference kernels. Versions 80 / 84 / 89."""
        from .train import Trainer
        return Trainer(self).train(examples)


class SantoriniNNetWrapper(NNetWrapper):
    """santorini/NNet.py:NNetWrapper (inference surface) for the no-god game. nn_args['nn_version'] must be 89."""
    NN_VERSION = 89

    def __init__(self, game, nn_args=None, state_dict=None, seed=0):
        nn_args = dict(nn_args or {'nn_version': 89})
        if nn_args.get('nn_version', 89) != 89:
            raise NotImplementedError('only SantoriniNNet version 89 is built (the shipped no-god checkpoint)')
        self.args = nn_args
        self.game = game
        self.board_size = game.getBoardSize(); self.action_size = game.getActionSize(); self.num_players = game.num_players
        self.requestKnowledgeTransfer = False
        self.state_dict = state_dict if state_dict is not None else random_v89_state_dict(seed)
        self.net = CudaNet(_lib.AZG_NET_SANTORINI_V89, game, v89_blob(self.state_dict))

    def load_state_dict(self, state_dict):
        self.state_dict = state_dict
        self.net.load(v89_blob(state_dict))


class AbaloneNNetWrapper(NNetWrapper):
    """abalone/NNet.py:NNetWrapper (inference surface). nn_args['nn_version'] must be 21."""
    NN_VERSION = 21

    def __init__(self, game, nn_args=None, state_dict=None, seed=0):
        nn_args = dict(nn_args or {'nn_version': 21})
        if nn_args.get('nn_version', 21) != 21:
            raise NotImplementedError('only AbaloneNNet version 21 is built (the shipped Belgian-daisy checkpoint)')
        self.args = nn_args
        self.game = game
        self.board_size = game.getBoardSize(); self.action_size = game.getActionSize(); self.num_players = game.num_players
        self.requestKnowledgeTransfer = False
        self.state_dict = state_dict if state_dict is not None else random_v21_state_dict(seed)
        self.net = CudaNet(_lib.AZG_NET_ABALONE_V21, game, v21_blob(self.state_dict))

    def load_state_dict(self, state_dict):
        self.state_dict = state_dict
        self.net.load(v21_blob(state_dict))


class AzulNNetWrapper(NNetWrapper):
    """azul/NNet.py:NNetWrapper (inference surface). nn_args['nn_version'] must be 84 (the shipped 2-player net)."""
    NN_VERSION = 84

    def __init__(self, game, nn_args=None, state_dict=None, seed=0):
        nn_args = dict(nn_args or {'nn_version': 84})
        if nn_args.get('nn_version', 84) != 84:
            raise NotImplementedError('only AzulNNet version 84 is built (the shipped 2-player checkpoint)')
        self.args = nn_args
        self.game = game
        self.board_size = game.getBoardSize(); self.action_size = game.getActionSize(); self.num_players = game.num_players
        self.requestKnowledgeTransfer = False
        self.state_dict = state_dict if state_dict is not None else random_v84_state_dict(seed)
        self.net = CudaNet(_lib.AZG_NET_AZUL_V84, game, v84_blob(self.state_dict))

    def load_state_dict(self, state_dict):
        self.state_dict = state_dict
        self.net.load(v84_blob(state_dict))


class HashNetWrapper:
    """Deterministic test net (tests only): prior/value are a hash of the board, see oracle/hashnet.py."""

    def __init__(self, game):
        self.game = game
        self.num_players = game.num_players
        self.net = CudaNet(_lib.AZG_NET_HASH, game, None)

    def predict(self, board, valid_actions):
        pi, v = self.net.forward(np.asarray(board)[None], np.asarray(valid_actions)[None])
        return pi[0], v[0]

    def predict_batch(self, boards, valid_actions):
        return self.net.forward(boards, valid_actions)
