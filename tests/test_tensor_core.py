"""Tensor-core matmul mode (tcgen05 kind::tf32): parity within the north star's 5e-3 relative bound.

On the fake device these tests exercise the host logic that only the tensor-core mode uses (padded
leading dimensions for TMA alignment, strided cross-entropy, un-copied transposed operands); on the
GPU they are the TF32 parity tests proper.
"""
import itertools
import numpy as np
import pytest
import lightgrad_b200 as light
import lightgrad_b200.nn as nn
from lightgrad_b200 import CudaTensor
from lightgrad_b200.autograd.cuda import ops
from tests import replay

TOL = 5e-3     # BASELINE.json north_star: TF32/BF16 tensor-core matmul and end-to-end gradients


@pytest.fixture(params=["fake", pytest.param("gpu", marks=pytest.mark.gpu)])
def device(request):
    request.getfixturevalue("fake_device" if request.param == "fake" else "cuda")
    prev = ops.set_matmul_mode('tf32')
    yield request.param
    ops.set_matmul_mode(prev)


def rel(got, want):
    return float(np.abs(got - want).max() / (np.abs(want).max() + 1e-30))


@pytest.mark.parametrize("shape", [(128, 256, 64), (300, 520, 136), (512, 70, 96), (768, 768, 2048), (130, 1026, 72)])
def test_matmul_all_layouts(device, shape):
    M, N, K = shape
    rs = np.random.RandomState(0)
    for ta, tb in itertools.product((False, True), repeat=2):
        a = rs.uniform(-1, 1, (K, M) if ta else (M, K)).astype(np.float32)
        b = rs.uniform(-1, 1, (N, K) if tb else (K, N)).astype(np.float32)
        A, B = CudaTensor.from_numpy(a), CudaTensor.from_numpy(b)
        out = (A.transpose(1, 0) if ta else A) @ (B.transpose(1, 0) if tb else B)
        want = (a.T if ta else a).astype(np.float64) @ (b.T if tb else b).astype(np.float64)
        assert out.shape == want.shape
        assert rel(out.numpy(), want) <= TOL, (shape, ta, tb)
        w = rs.uniform(-1, 1, want.shape).astype(np.float32)
        (out * CudaTensor.from_numpy(w)).sum().backward()
        ga = w.astype(np.float64) @ (b.T if tb else b).astype(np.float64).T
        gb = (a.T if ta else a).astype(np.float64).T @ w.astype(np.float64)
        assert rel(A.grad.numpy(), ga.T if ta else ga) <= TOL, (shape, ta, tb, 'dA')
        assert rel(B.grad.numpy(), gb.T if tb else gb) <= TOL, (shape, ta, tb, 'dB')


def test_linear_padded_vocab_and_cross_entropy(device):
    # out_features not a multiple of 4: the GEMM writes a padded buffer, cross entropy reads it by pitch
    rs = np.random.RandomState(1)
    x = rs.uniform(-1, 1, (4, 16, 64)).astype(np.float32)
    w = rs.uniform(-1, 1, (1022, 64)).astype(np.float32)
    b = rs.uniform(-1, 1, (1022,)).astype(np.float32)
    labels = rs.randint(0, 1022, size=(64,)).astype(np.int32)
    X, W, B = (CudaTensor.from_numpy(v) for v in (x, w, b))
    logits = X.linear(W, B)
    assert logits.shape == (4, 16, 1022)
    ref = x.astype(np.float64) @ w.T.astype(np.float64) + b
    assert rel(logits.numpy(), ref) <= TOL
    loss = light.loss.cross_entropy(logits.reshape(-1, 1022), CudaTensor.from_numpy(labels, requires_grad=False))
    loss.backward()
    z = ref.reshape(64, 1022)
    p = np.exp(z - z.max(axis=1, keepdims=True))
    p /= p.sum(axis=1, keepdims=True)
    want_loss = -np.log(p[np.arange(64), labels]).mean()
    assert abs(loss.item() - want_loss) <= TOL * abs(want_loss)
    p[np.arange(64), labels] -= 1
    dz = p / 64
    assert rel(W.grad.numpy(), dz.T @ x.reshape(64, 64).astype(np.float64)) <= TOL
    assert rel(B.grad.numpy(), dz.sum(axis=0)) <= TOL
    assert rel(X.grad.numpy(), (dz @ w.astype(np.float64)).reshape(x.shape)) <= TOL


def test_bert_tiny_tf32_matches_reference_golden(device):
    gmax = None
    items = list(replay.replay_bert_tiny(CudaTensor))
    gmax = max(float(np.abs(w).max()) for n, g, w in items if n.startswith('grad/'))
    for name, got, want in items:
        denom = max(float(np.abs(want).max()), 1e-4 * gmax)   # key.bias has a mathematically zero gradient
        assert float(np.abs(got - want).max()) / denom <= TOL, name
