"""Pins the oracle's V80 forward to the reference's SplendorNNet (torch CPU fp32) outputs
(tests/golden/splendor_v80_{rand,shipped}.npz). Tolerance 1e-5 absolute on pi and v (BASELINE.json)."""
import numpy as np

from oracle import oracle as O


def test_v80_forward_matches_reference(v80_golden):
    for tag, g in v80_golden.items():
        blob = O.v80_blob(g['sd'])
        assert blob.size == 142406 + 5 * 0 + sum(g['sd'][k].size for k in g['sd'] if 'running' in k)
        pi, v = O.v80_forward(blob, g['boards'], g['valids'])
        np.testing.assert_allclose(pi, g['pi'], rtol=0, atol=1e-5, err_msg=tag)
        np.testing.assert_allclose(v, g['v'], rtol=0, atol=1e-5, err_msg=tag)
        assert (pi[~g['valids']] == 0).all()


def test_v89_forward_matches_reference(v89_golden):
    """SantoriniNNet V89 (santorini/SantoriniNNet.py:194-217,273-279): oracle vs the reference's torch CPU fp32 outputs."""
    for tag, g in v89_golden.items():
        blob = O.v89_blob(g['sd'])
        assert blob.size == 381454 + sum(g['sd'][k].size for k in g['sd'] if 'running' in k)
        pi, v = O.v89_forward(blob, g['boards'], g['valids'])
        np.testing.assert_allclose(pi, g['pi'], rtol=0, atol=1e-5, err_msg=tag)
        np.testing.assert_allclose(v, g['v'], rtol=0, atol=1e-5, err_msg=tag)
        assert (pi[~g['valids']] == 0).all()


def test_v21_forward_matches_reference(v21_golden):
    """AbaloneNNet V21 (abalone/AbaloneNNet.py:117-156,173-202): oracle vs the reference's torch CPU fp32 outputs."""
    for tag, g in v21_golden.items():
        blob = O.v21_blob(g['sd'])
        assert blob.size == 35862 + sum(g['sd'][k].size for k in g['sd'] if 'running' in k)
        pi, v = O.v21_forward(blob, g['boards'], g['valids'])
        np.testing.assert_allclose(pi, g['pi'], rtol=0, atol=1e-5, err_msg=tag)
        np.testing.assert_allclose(v, g['v'], rtol=0, atol=1e-5, err_msg=tag)
        assert (pi[~g['valids']] == 0).all()
