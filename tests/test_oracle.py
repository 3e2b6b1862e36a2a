"""Pins the oracle (oracle/, a numpy restatement) and the host-side autograd core / nn / loss / optim /
example models against vectors produced by the REAL reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
from oracle import CpuTensor
from tests import replay

# same arithmetic in the same order on the same BLAS: agreement is to rounding of reassociated sums
RTOL, ATOL = 2e-6, 2e-6


def test_oracle_ops_match_reference_golden():
    n = 0
    for case, field, got, want in replay.replay_ops(CpuTensor):
        assert got.shape == want.shape, "%s/%s shape %s != %s" % (case, field, got.shape, want.shape)
        np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL, err_msg="%s/%s" % (case, field))
        n += 1
    assert n > 200


def test_oracle_bert_tiny_matches_reference_golden():
    n = 0
    for name, got, want in replay.replay_bert_tiny(CpuTensor):
        assert got.shape == want.shape, name
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6, err_msg=name)
        n += 1
    assert n > 40


def test_duplicate_index_gradient_is_accumulated():
    # the shipped reference assigns (last write wins); oracle and cuda backend add
    w = CpuTensor.from_numpy(np.arange(12, dtype=np.float32).reshape(4, 3))
    w[np.array([1, 1, 2])].sum().backward()
    np.testing.assert_array_equal(w.grad.numpy(), np.array([[0] * 3, [2] * 3, [1] * 3, [0] * 3], dtype=np.float32))


def test_walker_visits_shared_nodes_once():
    # h feeds two consumers; d/dx [exp(h) + h] with h = tanh(x)
    x = CpuTensor.from_numpy(np.linspace(-1, 1, 7).astype(np.float32))
    h = x.tanh()
    (h.exp() + h).sum().backward()
    hn = np.tanh(np.linspace(-1, 1, 7))
    np.testing.assert_allclose(x.grad.numpy(), (np.exp(hn) + 1) * (1 - hn ** 2), rtol=1e-5)


def test_leaf_hook_fires_after_last_contribution():
    # the data-parallel wrapper relies on this to start reducing finished gradients during backward
    from lightgrad_b200.autograd import Gradients
    a = CpuTensor.from_numpy(np.ones((3, 3), dtype=np.float32))
    b = CpuTensor.from_numpy(np.ones((3, 3), dtype=np.float32))
    seen = []

    def hook(t):
        seen.append((t is a, t.grad.numpy().copy()))
    y = ((a * b) + a.exp() + a).sum()        # a has three consumers, b one
    a.zero_grad(), b.zero_grad()
    Gradients.leaf_hook = hook
    try:
        y.backward()
    finally:
        Gradients.leaf_hook = None
    assert len(seen) == 2
    for is_a, g in seen:
        want = (1 + np.e + 1) if is_a else 1.0
        np.testing.assert_allclose(g, want, rtol=1e-6)      # complete at the time the hook ran


def test_leaf_hook_waits_for_wrapper_nodes_that_reach_the_leaf_by_closure():
    # a WrapperFunction back-propagates through a private graph; a parameter it uses by closure receives a contribution
    # the outer walk cannot count, so no leaf is reported as finished while such a node is still outstanding
    from lightgrad_b200.autograd import Gradients, WrapperFunction
    w = CpuTensor.from_numpy(np.full((3,), 2.0, dtype=np.float32))
    x = CpuTensor.from_numpy(np.arange(3, dtype=np.float32))

    @WrapperFunction.from_function
    def scaled(t):
        return t * w                       # w reached by closure, not as a parent of the wrapper node

    seen = {}

    def hook(t):
        seen[id(t)] = t.grad.numpy().copy()
    y = (scaled(x).sum() + (w * 3.0).sum())          # w is ALSO a direct parent of an outer node (processed first)
    w.zero_grad(), x.zero_grad()
    Gradients.leaf_hook = hook
    try:
        y.backward()
    finally:
        Gradients.leaf_hook = None
    # at hook time w already held both contributions: 3 (direct) + x (through the wrapper's private graph)
    np.testing.assert_allclose(seen[id(w)], 3.0 + np.arange(3), rtol=1e-6)
    np.testing.assert_allclose(w.grad.numpy(), 3.0 + np.arange(3), rtol=1e-6)
