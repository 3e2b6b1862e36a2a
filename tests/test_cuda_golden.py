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


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2048, 1280, 144), (1280, 2048, 272), (4, 1024, 768, 64)])
def test_exact_fp32_matmul_vectorised_path_all_layouts(cuda, shape):
    """The exact-fp32 SIMT GEMM's fast path (whole 128x128x16 tiles, 128-bit loads, swizzled staging) in all four operand
    layouts, with a bias, as a batch and accumulating into an existing gradient -- against float64 numpy at <= 1e-5."""
    import itertools
    from lightgrad_b200.autograd.cuda import ops
    prev = ops.set_matmul_mode('fp32')
    try:
        rs = np.random.RandomState(3)
        batch = shape[:-3]
        M, N, K = shape[-3:]
        for ta, tb in itertools.product((False, True), repeat=2):
            a = rs.uniform(-1, 1, batch + ((K, M) if ta else (M, K))).astype(np.float32)
            b = rs.uniform(-1, 1, (N, K) if tb else (K, N)).astype(np.float32)
            A, B = CudaTensor.from_numpy(a), CudaTensor.from_numpy(b)
            nd = len(a.shape)
            perm = tuple(range(nd - 2)) + (nd - 1, nd - 2)
            out = (A.transpose(*perm) if ta else A) @ (B.transpose(1, 0) if tb else B)
            a64 = (np.swapaxes(a, -1, -2) if ta else a).astype(np.float64)
            b64 = (b.T if tb else b).astype(np.float64)
            want = a64 @ b64
            err = np.abs(out.numpy() - want).max() / np.abs(want).max()
            assert err <= 1e-5, (shape, ta, tb, err)
            w = rs.uniform(-1, 1, want.shape).astype(np.float32)
            (out * CudaTensor.from_numpy(w)).sum().backward()
            # a second product of the same operands: its backward ACCUMULATES into the existing .grad
            out2 = (A.transpose(*perm) if ta else A) @ (B.transpose(1, 0) if tb else B)
            (out2 * CudaTensor.from_numpy(w)).sum().backward()
            ga = 2 * (w.astype(np.float64) @ b64.T)
            gb = 2 * (a64.reshape(-1, K).T @ w.astype(np.float64).reshape(-1, N))
            ga = np.swapaxes(ga, -1, -2) if ta else ga
            gb = gb.T if tb else gb
            assert np.abs(A.grad.numpy() - ga).max() / np.abs(ga).max() <= 1e-5, (shape, ta, tb, 'dA')
            assert np.abs(B.grad.numpy() - gb).max() / np.abs(gb).max() <= 1e-5, (shape, ta, tb, 'dB')
        # linear layer with a bias (K-contiguous weight, bias added in the kernel)
        x = rs.uniform(-1, 1, (2048, 256)).astype(np.float32)
        wt = rs.uniform(-1, 1, (1280, 256)).astype(np.float32)
        bs = rs.uniform(-1, 1, (1280,)).astype(np.float32)
        y = CudaTensor.from_numpy(x).linear(CudaTensor.from_numpy(wt), CudaTensor.from_numpy(bs))
        want = x.astype(np.float64) @ wt.astype(np.float64).T + bs
        assert np.abs(y.numpy() - want).max() / np.abs(want).max() <= 1e-5
    finally:
        ops.set_matmul_mode(prev)
