"""Training step for the nets of the self-play path (reference: GenericNNetWrapper.train / loss_pi / loss_v,
GenericNNetWrapper.py:44-92,179-190), fed from DEVICE-RESIDENT examples: the tensors `Engine.examples_device()` drains from the
self-play ring (or `dist.gather_examples` returns on the training rank) go straight into the optimiser loop -- no pickle, no host copy.

SURVEY 8f-1 ("next" row). The forward / backward math here is PyTorch (autograd, cuBLAS/cuDNN kernels): library code, NOT one of the
hand-written sm_100a kernels -- those cover the inference path that dominates self-play. What this module guarantees is the
reference's training SEMANTICS and weight interchange:
  * modules named like the reference's (splendor/SplendorNNet.py:149-204,259-280; azul/AzulNNet.py:84-111;
    santorini/SantoriniNNet.py:16-40,70-84,194-217), so `state_dict()` keys and shapes are the reference's: weights move both ways
    through `nnet.load_state_dict` / `formats.save_checkpoint_file`, and the CUDA inference kernels load the trained weights directly
  * the same loop: AdamW(lr=learn_rate), OneCycleLR(max_lr=learn_rate, steps_per_epoch=len/batch, epochs), per batch
    `np.random.choice(len, batch_size, replace=False)`, loss = KLDiv(batchmean)(log pi, target pi) + 0.25 x loss_v with the Q/Z mix
    `(z + q_weight q) / (1 + q_weight)` normalised by batch x players (GenericNNetWrapper.py:60-79,179-190)
tests/test_train.py checks one optimiser step against the reference's own `train` on the same batch (losses and every updated
tensor), in the build container where the reference runs, and against a recorded fixture elsewhere.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def _make_divisible(v, divisor, min_value=None):
    min_value = divisor if min_value is None else min_value
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


class LinearNormActivation(nn.Module):
    """Token-axis Linear (or feature-axis Linear when depthwise) + BatchNorm1d + activation, SplendorNNet.py:149-170."""

    def __init__(self, in_size, out_size, activation_layer, depthwise=False, channels=None):
        super().__init__()
        self.linear = nn.Linear(in_size, out_size, bias=False)
        self.norm = nn.BatchNorm1d(channels if depthwise else out_size)
        self.activation = activation_layer(inplace=True) if activation_layer is not None else nn.Identity()
        self.depthwise = depthwise

    def forward(self, x):
        y = self.linear(x) if self.depthwise else self.linear(x.transpose(-1, -2)).transpose(-1, -2)
        return self.activation(self.norm(y))


class SqueezeExcitation1d(nn.Module):
    """SplendorNNet.py:172-187."""

    def __init__(self, input_channels, squeeze_channels, setype='avg'):
        super().__init__()
        self.avgpool = nn.AdaptiveAvgPool1d(1) if setype == 'avg' else nn.AdaptiveMaxPool1d(1)
        self.fc1 = nn.Linear(input_channels, squeeze_channels)
        self.activation = nn.ReLU()
        self.fc2 = nn.Linear(squeeze_channels, input_channels)
        self.scale_activation = nn.Hardsigmoid()

    def forward(self, x):
        s = self.avgpool(x)
        s = self.activation(self.fc1(s.transpose(-1, -2)).transpose(-1, -2))
        s = self.fc2(s.transpose(-1, -2)).transpose(-1, -2)
        return self.scale_activation(s) * x


class InvertedResidual1d(nn.Module):
    """SplendorNNet.py:189-204 (use_se is always truthy in the reference's V80 / V84 constructors)."""

    def __init__(self, in_channels, exp_channels, out_channels, kernel, use_hs, setype='avg'):
        super().__init__()
        self.use_res_connect = in_channels == out_channels
        act = nn.Hardswish if use_hs else nn.ReLU
        self.expand = LinearNormActivation(in_channels, exp_channels, act)
        self.depthwise = LinearNormActivation(kernel, kernel, act, depthwise=True, channels=exp_channels)
        self.se = SqueezeExcitation1d(exp_channels, _make_divisible(exp_channels // 4, 8), setype)
        self.project = LinearNormActivation(exp_channels, out_channels, None)

    def forward(self, x):
        y = self.project(self.se(self.depthwise(self.expand(x))))
        return y + x if self.use_res_connect else y


class TokenMixerNet(nn.Module):
    """SplendorNNet version 80 (56 x 7 boards, 2 players) and AzulNNet version 84 (23 x 6 boards)."""

    def __init__(self, version, nb_vect, vect_dim, action_size, num_players, dropout=0.0):
        super().__init__()
        self.version, self.nb_vect, self.vect_dim, self.dropout = version, nb_vect, vect_dim, dropout
        nv = nb_vect
        if version == 80:
            trunk = InvertedResidual1d(nv, 3 * nv, nv, vect_dim, False)
            pi_blk = InvertedResidual1d(nv, 3 * nv, nv, vect_dim, True, setype='max'); pi_flat = nv * vect_dim
            v_blk = InvertedResidual1d(nv, 3 * nv, nv, vect_dim, True, setype='max')
        elif version == 84:
            trunk = InvertedResidual1d(nv, 5 * nv, nv, vect_dim, False)
            pi_blk = InvertedResidual1d(nv, 5 * nv, 2 * nv, vect_dim, True, setype='avg'); pi_flat = 2 * nv * vect_dim
            v_blk = InvertedResidual1d(nv, 2 * nv, nv, vect_dim, True, setype='avg')
        else:
            raise NotImplementedError(version)
        self.first_layer = LinearNormActivation(nv, nv, None)
        self.trunk = nn.Sequential(trunk)
        self.output_layers_PI = nn.Sequential(pi_blk, nn.Flatten(1), nn.Linear(pi_flat, action_size), nn.ReLU(), nn.Linear(action_size, action_size))
        self.output_layers_V = nn.Sequential(v_blk, nn.Flatten(1), nn.Linear(nv * vect_dim, num_players), nn.ReLU(), nn.Linear(num_players, num_players))
        self.register_buffer('lowvalue', torch.FloatTensor([-1e8]))

    def forward(self, boards, valid_actions):
        x = boards.view(-1, self.nb_vect, self.vect_dim)
        x = self.first_layer(x)
        x = F.dropout(self.trunk(x), p=self.dropout, training=self.training)
        v = self.output_layers_V(x)
        pi = torch.where(valid_actions, self.output_layers_PI(x), self.lowvalue)
        return F.log_softmax(pi, dim=1), torch.tanh(v)


class _ResBlock(nn.Module):
    """SimpleResBlock, santorini/SantoriniNNet.py:70-84."""

    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.Conv2d(c, c, 3, padding=1, bias=False); self.bn1 = nn.BatchNorm2d(c)
        self.conv2 = nn.Conv2d(c, c, 3, padding=1, bias=False); self.bn2 = nn.BatchNorm2d(c)

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        return F.relu(self.bn2(self.conv2(y)) + x)


class _HeadPI(nn.Module):
    def __init__(self, c, n_out):
        super().__init__()
        self.conv1x1 = nn.Conv2d(c, 2, 1, bias=False); self.bn = nn.BatchNorm2d(2); self.fc = nn.Linear(50, n_out)

    def forward(self, x):
        return self.fc(torch.flatten(F.relu(self.bn(self.conv1x1(x))), 1))


class _HeadV(nn.Module):
    def __init__(self, c, n_players):
        super().__init__()
        self.conv1x1 = nn.Conv2d(c, 1, 1, bias=False); self.bn = nn.BatchNorm2d(1); self.fc1 = nn.Linear(25, 64); self.fc2 = nn.Linear(64, n_players)

    def forward(self, x):
        return self.fc2(F.relu(self.fc1(torch.flatten(F.relu(self.bn(self.conv1x1(x))), 1))))


class SantoriniV89Net(nn.Module):
    """SantoriniNNet version 89 without gods (santorini/SantoriniNNet.py:194-217,273-279): 5x5x3 boards, channels 0-1 into a
    64-wide 3x3 residual trunk of 5 blocks, conv1x1 heads."""

    def __init__(self, action_size=162, num_players=2):
        super().__init__()
        self.first_layer = nn.Sequential(nn.Conv2d(2, 64, 3, padding=1, bias=False), nn.BatchNorm2d(64), nn.ReLU())
        self.trunk = nn.Sequential(*[_ResBlock(64) for _ in range(5)])
        self.head_PI = _HeadPI(64, action_size); self.head_V = _HeadV(64, num_players)
        self.register_buffer('lowvalue', torch.FloatTensor([-1e8]))

    def forward(self, boards, valid_actions):
        x = boards.view(-1, 5, 5, 3).permute(0, 3, 1, 2)[:, :2].contiguous()
        x = self.trunk(self.first_layer(x))
        pi = torch.where(valid_actions, self.head_PI(x), self.lowvalue)
        return F.log_softmax(pi, dim=1), torch.tanh(self.head_V(x))


def build_net(kind, game):
    """torch module for net kind 80 / 84 / 89 with the reference's parameter names."""
    if kind in (80, 84):
        nv, d = game.getBoardSize()[:2]
        return TokenMixerNet(kind, nv, d, game.getActionSize(), game.num_players)
    if kind == 89:
        return SantoriniV89Net(game.getActionSize(), game.num_players)
    raise NotImplementedError(f'no training module for net version {kind} (built: 80, 84, 89)')


def loss_pi(targets, outputs):
    """GenericNNetWrapper.loss_pi (:179-181): KL divergence, batchmean, outputs = log-probabilities."""
    return F.kl_div(outputs, targets, reduction='batchmean')


def loss_v(targets_v, targets_q, outputs, q_weight):
    """GenericNNetWrapper.loss_v (:188-190)."""
    targets = (targets_v + q_weight * targets_q) / (1 + q_weight)
    return torch.sum((targets - outputs) ** 2) / (targets_v.size()[0] * targets_v.size()[-1])


class Trainer:
    """GenericNNetWrapper.train (:44-92) on one GPU (or CPU) for a wrapper of this package: after train() the wrapper's CUDA
    inference net holds the new weights."""

    def __init__(self, wrapper, args=None, device=None):
        self.wrapper = wrapper
        self.args = dict(learn_rate=3e-4, epochs=2, batch_size=32, q_weight=0.5, dropout=0.0)
        self.args.update({k: v for k, v in dict(wrapper.args).items() if k in self.args})
        self.args.update(args or {})
        self.device = torch.device(device if device is not None else ('cuda' if torch.cuda.is_available() else 'cpu'))
        self.model = build_net(wrapper.NN_VERSION, wrapper.game).to(self.device)
        if hasattr(self.model, 'dropout'):
            self.model.dropout = float(self.args['dropout'])
        self.load_from_wrapper()

    def load_from_wrapper(self):
        sd = {k: torch.as_tensor(np.asarray(v)) for k, v in self.wrapper.state_dict.items()}
        missing, unexpected = self.model.load_state_dict(sd, strict=False)
        bad = [k for k in missing if not k.endswith('num_batches_tracked') and k != 'lowvalue']
        if bad or unexpected:
            raise KeyError(f'state_dict mismatch: missing {bad}, unexpected {list(unexpected)}')

    def push_to_wrapper(self):
        """New weights -> the wrapper's state_dict and its CUDA inference kernels (azg_net_load)."""
        sd = {k: v.detach().cpu().numpy() for k, v in self.model.state_dict().items()}
        self.wrapper.load_state_dict(sd)

    def _as_tensors(self, examples):
        """examples: (boards, pi, z, valids, q) arrays / tensors, or a list / deque of the reference's example tuples."""
        if not (isinstance(examples, tuple) and len(examples) == 5 and hasattr(examples[0], 'shape')):
            from .formats import examples_to_arrays
            examples = examples_to_arrays(examples)
        b, pi, z, va, q = examples
        to = lambda x, dt: (x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))).to(self.device).to(dt)
        return to(b, torch.float32).reshape(len(b), -1), to(pi, torch.float32), to(z, torch.float32), to(va, torch.bool), to(q, torch.float32)

    def train_step(self, optimizer, batch):
        boards, target_pis, target_vs, valids, target_qs = batch
        optimizer.zero_grad(set_to_none=True)
        out_pi, out_v = self.model(boards, valids)
        l_pi, l_v = loss_pi(target_pis, out_pi), loss_v(target_vs, target_qs, out_v, self.args['q_weight'])
        (l_pi + 0.25 * l_v).backward()                            # total_loss = l_pi + 0.25 * l_v, GenericNNetWrapper.py:72-80
        optimizer.step()
        return float(l_pi.item()), float(l_v.item())

    def train(self, examples, sample_ids=None):
        """The reference's loop. sample_ids (tests): explicit [n_batches][batch_size] indices instead of np.random.choice."""
        a = self.args
        data = self._as_tensors(examples)
        n = len(data[0])
        batch_count = int(n / a['batch_size'])
        optimizer = torch.optim.AdamW(self.model.parameters(), lr=a['learn_rate'])
        scheduler = torch.optim.lr_scheduler.OneCycleLR(optimizer, max_lr=a['learn_rate'], steps_per_epoch=batch_count, epochs=a['epochs'])
        losses = []
        step = 0
        for epoch in range(a['epochs']):
            self.model.train()
            for _ in range(batch_count):
                ids = np.random.choice(n, size=a['batch_size'], replace=False) if sample_ids is None else np.asarray(sample_ids[step])
                idx = torch.as_tensor(ids, device=self.device, dtype=torch.long)
                losses.append(self.train_step(optimizer, tuple(t.index_select(0, idx) for t in data)))
                scheduler.step()
                step += 1
        self.model.eval()
        self.push_to_wrapper()
        return losses
