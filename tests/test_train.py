"""Training step: Trainer.train against the reference's OWN GenericNNetWrapper.train on the same examples and the same
sample ids (tests/golden/splendor_train_step.npz, recorded by oracle/gen_golden_train.py running the unmodified reference):
per-batch policy / value losses and every tensor of the state_dict after three AdamW + OneCycleLR steps. CPU here; the GPU variant
(-m gpu) also pushes the trained weights into the CUDA inference kernel and checks its outputs against the torch module."""
import os

import numpy as np
import pytest
import torch

from azg_b200.train import Trainer, TokenMixerNet, SantoriniV89Net, build_net
from conftest import GOLDEN


class FakeGame:
    num_players = 2
    def __init__(self, shape, actions): self.shape, self.actions = shape, actions
    def getBoardSize(self): return self.shape
    def getActionSize(self): return self.actions


class FakeWrapper:
    def __init__(self, version, game, sd, args):
        self.NN_VERSION = version; self.game = game; self.state_dict = sd; self.args = args; self.loaded = None
    def load_state_dict(self, sd): self.state_dict = sd; self.loaded = sd


def _golden():
    z = np.load(os.path.join(GOLDEN, 'splendor_train_step.npz'))
    sd0 = {k[5:]: z[k] for k in z.files if k.startswith('sd0__')}; sd1 = {k[5:]: z[k] for k in z.files if k.startswith('sd1__')}
    return z, sd0, sd1


def _check(tr, z, sd1, atol):
    losses = tr.train((z['boards'], z['pi'], z['z'], z['valids'], z['q']), sample_ids=z['ids'])
    np.testing.assert_allclose(np.array(losses), z['losses'], rtol=2e-5, atol=1e-6)
    got = tr.wrapper.loaded
    assert set(got) == set(sd1)
    for k in sd1:
        if k.endswith('num_batches_tracked'):
            assert int(got[k]) == int(sd1[k]) == 3
        else:
            np.testing.assert_allclose(got[k], sd1[k], rtol=0, atol=atol, err_msg=k)


def test_three_optimizer_steps_match_the_reference_cpu():
    torch.set_num_threads(1)
    z, sd0, sd1 = _golden()
    w = FakeWrapper(80, FakeGame((56, 7), 81), sd0, dict(learn_rate=1e-3, epochs=1, batch_size=16, q_weight=0.5, dropout=0.0))
    _check(Trainer(w, device='cpu'), z, sd1, atol=2e-6)
    changed = [k for k in sd0 if not k.endswith('num_batches_tracked') and k != 'lowvalue' and not np.array_equal(sd0[k], sd1[k])]
    assert len(changed) > 60                                     # the step really moved (nearly) every tensor


def test_module_names_and_shapes_follow_the_reference():
    from azg_b200.nnet import V80_TENSOR_ORDER, V84_TENSOR_ORDER, V89_TENSOR_ORDER, random_v84_state_dict, random_v89_state_dict
    for ver, game, order, rnd in ((84, FakeGame((23, 6), 180), V84_TENSOR_ORDER, random_v84_state_dict), (89, FakeGame((5, 5, 3), 162), V89_TENSOR_ORDER, random_v89_state_dict)):
        m = build_net(ver, game); sd = m.state_dict(); r = rnd(0)
        assert all(k in sd and tuple(sd[k].shape) == r[k].shape for k in order), ver
    z, sd0, _ = _golden()
    sd = build_net(80, FakeGame((56, 7), 81)).state_dict()
    assert set(sd) == set(sd0) and all(tuple(sd[k].shape) == sd0[k].shape for k in sd0) and all(k in sd for k in V80_TENSOR_ORDER)


@pytest.mark.gpu
def test_train_on_device_examples_and_push_to_the_cuda_net():
    """The whole loop on one GPU: self-play examples stay in HBM (examples_device) -> Trainer -> azg_net_load; the CUDA inference
    kernel then agrees with the trained torch module to 1e-5, and the reference-recorded step is reproduced on the GPU too."""
    import azg_b200
    from azg_b200.mcts import Engine
    z, sd0, sd1 = _golden()
    game = azg_b200.SplendorGame()
    net = azg_b200.NNetWrapper(game, dict(nn_version=80, learn_rate=1e-3, epochs=1, batch_size=16, q_weight=0.5), state_dict=sd0)
    tr = Trainer(net, device='cuda')
    losses = tr.train((z['boards'], z['pi'], z['z'], z['valids'], z['q']), sample_ids=z['ids'])
    np.testing.assert_allclose(np.array(losses), z['losses'], rtol=1e-4, atol=1e-5)            # cuBLAS / cuDNN summation order differs from the CPU's
    # AdamW normalises every element's step by its own gradient history: an element whose gradient is ~0 moves by up to +-lr whatever
    # the gradient's size, so the summation-order noise of the GPU GEMMs can flip single elements by ~lr (1e-3 at the peak of the
    # one-cycle schedule). Bar: 99.9 % of all elements within 2e-5, none further than 2 lr.
    n_all = n_far = 0
    for k in sd1:
        if not k.endswith('num_batches_tracked'):
            d = np.abs(np.asarray(net.state_dict[k], np.float64) - sd1[k])
            n_all += d.size; n_far += int((d > 2e-5).sum())
            assert d.max() < 2e-3, k
    assert n_far <= 1e-3 * n_all, (n_far, n_all)
    b = z['boards'][:32]; va = z['valids'][:32]
    pi, v = net.predict_batch(b, va)
    with torch.no_grad():
        lp, tv = tr.model(torch.from_numpy(b.astype(np.float32)).cuda().reshape(32, -1), torch.from_numpy(va).cuda())
    assert np.abs(pi - torch.exp(lp).cpu().numpy()).max() < 1e-5 and np.abs(v - tv.cpu().numpy()).max() < 1e-5
    # self-play -> device-resident examples -> another round of training, nothing leaves the GPU
    eng = Engine(game, net, dict(numMCTSSims=16, universes=1, prob_fullMCTS=1.0), n_games=64, dirichlet_noise=True, seed=1, node_cap=256)
    eng.selfplay(min_episodes=8)
    ex = eng.examples_device(); eng.close()
    assert ex[0].is_cuda and len(ex[0]) >= 32
    tr2 = Trainer(net, args=dict(batch_size=32, epochs=1), device='cuda')
    l2 = tr2.train(ex)
    assert len(l2) == len(ex[0]) // 32 and all(np.isfinite(l2).ravel())
