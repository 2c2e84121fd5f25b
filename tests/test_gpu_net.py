"""GPU parity of the V80 forward kernel against the reference's SplendorNNet outputs (golden, torch CPU fp32) and
against the CPU oracle on a larger batch. Tolerance 1e-5 absolute on pi and v (BASELINE.json north_star)."""
import numpy as np
import pytest

import azg_b200
from azg_b200.nnet import NNetWrapper, HashNetWrapper
from oracle import oracle as O
from oracle.hashnet import hashnet_eval

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope='module')
def game():
    return azg_b200.SplendorGame()


@pytest.mark.parametrize('tag', ['rand', 'shipped'])
def test_v80_forward_golden(game, v80_golden, tag):
    g = v80_golden[tag]
    net = NNetWrapper(game, {'nn_version': 80}, state_dict=g['sd'])
    pi, v = net.predict_batch(g['boards'], g['valids'])
    np.testing.assert_allclose(pi, g['pi'], rtol=0, atol=TOL)
    np.testing.assert_allclose(v, g['v'], rtol=0, atol=TOL)
    assert (pi[~g['valids']] == 0).all()
    np.testing.assert_allclose(pi.sum(axis=1), 1.0, atol=1e-5)
    # scalar predict == row of the batch (NeuralNet.predict surface)
    p0, v0 = net.predict(g['boards'][3], g['valids'][3])
    assert p0.shape == (81,) and v0.shape == (2,) and p0.dtype == np.float32
    np.testing.assert_allclose(p0, g['pi'][3], rtol=0, atol=TOL)


def test_v80_forward_vs_oracle_ragged_batches(game, v80_golden, kat):
    sd = v80_golden['rand']['sd']
    net = NNetWrapper(game, {'nn_version': 80}, state_dict=sd)
    blob = O.v80_blob(sd)
    for n in (1, 15, 16, 17, 333):                          # partial tiles of the 16-leaf CTA tile
        b = kat['canonical'][:n]; va = kat['valids'][:n]
        pi, v = net.predict_batch(b, va)
        opi, ov = O.v80_forward(blob, b, va)
        np.testing.assert_allclose(pi, opi, rtol=0, atol=TOL)
        np.testing.assert_allclose(v, ov, rtol=0, atol=TOL)


def test_hashnet_bit_exact(game, kat):
    net = HashNetWrapper(game)
    b = kat['canonical'][::5]; va = kat['valids'][::5]
    pi, v = net.predict_batch(b, va)
    for i in range(len(b)):
        p2, v2 = hashnet_eval(b[i], va[i])
        assert (pi[i] == p2).all() and (v[i] == v2).all()


def test_v80_tensor_core_kernel_large_batch(game, v80_golden, kat, monkeypatch):
    """The tcgen05 (3xTF32) kernel against the fp32 CUDA-core kernel and the oracle on a batch that gives every persistent
    CTA several tiles (5000 leaves = 313 tiles > 148 SMs) with a ragged last tile; both weight sets."""
    reps = 9
    b = np.concatenate([kat['canonical']] * reps)[:5000]; va = np.concatenate([kat['valids']] * reps)[:5000]
    for tag in ('rand', 'shipped'):
        sd = v80_golden[tag]['sd']
        monkeypatch.delenv('AZG_V80_KERNEL', raising=False)
        net_tc = NNetWrapper(game, {'nn_version': 80}, state_dict=sd)
        monkeypatch.setenv('AZG_V80_KERNEL', 'fp32')
        net_f32 = NNetWrapper(game, {'nn_version': 80}, state_dict=sd)
        monkeypatch.delenv('AZG_V80_KERNEL', raising=False)
        pi, v = net_tc.predict_batch(b, va)
        pi2, v2 = net_f32.predict_batch(b, va)
        np.testing.assert_allclose(pi, pi2, rtol=0, atol=TOL)
        np.testing.assert_allclose(v, v2, rtol=0, atol=TOL)
        # identical boards in different tiles / CTAs give identical results (no cross-tile state)
        n0 = len(kat['canonical'])
        assert (pi[:n0] == pi[n0:2 * n0]).all() and (v[:n0] == v[n0:2 * n0]).all()
        opi, ov = O.v80_forward(O.v80_blob(sd), b[:64], va[:64])
        np.testing.assert_allclose(pi[:64], opi, rtol=0, atol=TOL)
        np.testing.assert_allclose(v[:64], ov, rtol=0, atol=TOL)
