"""BF16 tensor-core matmul mode (LG_GEMM_BF16_TC): tcgen05 kind::f16 on bf16 STAGING COPIES of the fp32 operands,
fp32 accumulation and results.

Oracle: the CPU tensor's matmul (numpy, cpu/ops.py:107-116 of the reference) applied to the operands rounded to bf16
the way lg_cast rounds them (nearest even) -- against that the kernel must agree to fp32 summation noise (<= 1e-5),
which pins the operand layouts (K-major / MN-major, chunked and unchunked maps, batched, ragged, CTA pairs);
against the unrounded product the stated bound for tensor-core modes is 5e-3.
Host logic (CPU suite, numpy test double): the staging copies are shared by all views of a buffer and dropped when it
is written."""
import numpy as np
import pytest
import lightgrad_b200 as light
from lightgrad_b200 import CudaTensor
from lightgrad_b200.autograd.cuda import ops, runtime as rt


def _bf16_round(a):
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) >> 16 << 16
    return u.astype(np.uint32).view(np.float32).reshape(a.shape)


@pytest.fixture
def bf16_mode():
    prev = ops.set_matmul_mode('bf16')
    yield
    ops.set_matmul_mode(prev)


def _check_product(a, b, ta=False, tb=False, tol_exact=2e-5):
    """a @ b with optionally transposed VIEWS as operands (exercises the MN-major layouts)."""
    A = CudaTensor.from_numpy(a.T.copy() if ta else a)
    B = CudaTensor.from_numpy(b.T.copy() if tb else b)
    with light.no_grad():
        Av = A.transpose() if ta else A
        Bv = B.transpose() if tb else B
        got = (Av @ Bv).numpy()
    want16 = _bf16_round(a).astype(np.float64) @ _bf16_round(b).astype(np.float64)
    want = a.astype(np.float64) @ b.astype(np.float64)
    scale = np.abs(want).max()
    # operand strides that are not 16-byte multiples of bf16 (a 30522-wide row-major B) cannot go through TMA as
    # bf16: that product runs exactly, on the fp32 data -- agreeing with the unrounded product instead
    e16, e32 = np.abs(got - want16).max() / scale, np.abs(got - want).max() / scale
    assert min(e16, e32) <= tol_exact, (a.shape, b.shape, ta, tb, e16, e32)
    assert e32 <= 5e-3


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(256, 256, 256), (128, 64, 512), (384, 200, 136), (1000, 520, 264), (4096, 768, 768),
                                   (512, 3072, 768), (300, 30522, 64)])
def test_bf16_matmul_all_operand_layouts(cuda, bf16_mode, shape):
    if not ops.matmul_mode_available('bf16'):
        pytest.skip("bf16 tensor-core mode is not built")
    M, N, K = shape
    rs = np.random.RandomState(M + N + K)
    a = rs.uniform(-1, 1, (M, K)).astype(np.float32)
    b = rs.uniform(-1, 1, (K, N)).astype(np.float32)
    for ta in (False, True):
        for tb in (False, True):
            _check_product(a, b, ta, tb)


@pytest.mark.gpu
def test_bf16_matmul_4096_cubed_pair_kernel_and_batched_heads(cuda, bf16_mode):
    if not ops.matmul_mode_available('bf16'):
        pytest.skip("bf16 tensor-core mode is not built")
    rs = np.random.RandomState(5)
    a = rs.uniform(-1, 1, (4096, 4096)).astype(np.float32)
    b = rs.uniform(-1, 1, (4096, 4096)).astype(np.float32)
    A, B = CudaTensor.from_numpy(a), CudaTensor.from_numpy(b)
    with light.no_grad():
        got = (A @ B).numpy()
    rows = rs.choice(4096, 24, replace=False)
    want16 = _bf16_round(a[rows]).astype(np.float64) @ _bf16_round(b).astype(np.float64)
    assert np.abs(got[rows] - want16).max() <= 2e-5 * np.abs(want16).max()
    # batched per-head views of one (rows, 3H) buffer, as the attention block issues them
    bsz, heads, s, dh = 4, 12, 128, 64
    H = heads * dh
    qkv = rs.uniform(-1, 1, (3, bsz * s, H)).astype(np.float32)
    Q = CudaTensor.from_numpy(qkv)
    hv = (bsz, heads, s, dh), (s * H, dh, H, 1)
    q, k = Q._view(hv[0], hv[1], 0), Q._view(hv[0], hv[1], bsz * s * H)
    with light.no_grad():
        scores = ops._gemm(q, ops._swap_last(k)).numpy()
    qn = _bf16_round(qkv[0]).reshape(bsz, s, heads, dh).transpose(0, 2, 1, 3).astype(np.float64)
    kn = _bf16_round(qkv[1]).reshape(bsz, s, heads, dh).transpose(0, 2, 1, 3).astype(np.float64)
    want = qn @ kn.transpose(0, 1, 3, 2)
    assert np.abs(scores - want).max() <= 2e-5 * np.abs(want).max()


@pytest.mark.gpu
def test_bf16_linear_backward_and_fused_nodes_against_the_oracle(cuda, bf16_mode):
    if not ops.matmul_mode_available('bf16'):
        pytest.skip("bf16 tensor-core mode is not built")
    from oracle import CpuTensor
    rs = np.random.RandomState(0)
    x = rs.uniform(-1, 1, (4, 128, 768)).astype(np.float32)
    w = (rs.uniform(-1, 1, (768, 768)) / 28).astype(np.float32)
    b = rs.uniform(-1, 1, (768,)).astype(np.float32)
    res = {}
    for T in (CpuTensor, CudaTensor):
        X, W, Bv = T.from_numpy(x), T.from_numpy(w), T.from_numpy(b)
        y = X.linear(W, Bv) if hasattr(X, 'linear') else X @ W.T(1, 0) + Bv
        y = y.gelu() if hasattr(y, 'gelu') else 0.5 * y * (1.0 + (y * 0.7978845608 * (1.0 + 0.044715 * y * y)).tanh())
        y.sum().backward()
        res[T.__name__] = (y.numpy(), X.grad.numpy(), W.grad.numpy(), Bv.grad.numpy())
    for want, got in zip(res['CpuTensor'], res['CudaTensor']):
        assert np.abs(want - got).max() <= 5e-3 * np.abs(want).max()


def test_staging_copies_follow_writes_host_logic(fake_device, bf16_mode):
    rs = np.random.RandomState(1)
    a = rs.uniform(-1, 1, (256, 64)).astype(np.float32)
    b = rs.uniform(-1, 1, (64, 256)).astype(np.float32)
    A, B = CudaTensor.from_numpy(a), CudaTensor.from_numpy(b)
    with light.no_grad():
        fake_device.mode_log = []
        c1 = (A @ B).numpy()
        assert fake_device.mode_log == [rt.GEMM_BF16_TC]
        assert A._data._bf16 is not None and B._data._bf16 is not None
        n_casts = fake_device.launches
        c1b = (A.transpose().transpose() @ B).numpy()             # views share the copy: no new conversion
        assert fake_device.launches == n_casts + 1
        np.testing.assert_array_equal(c1, c1b)
        A *= 2.0                                                   # in-place write drops the stale copy
        assert A._data._bf16 is None
        c2 = (A @ B).numpy()
    np.testing.assert_allclose(c2, 2 * c1, rtol=1e-6)
    np.testing.assert_allclose(c1, _bf16_round(a).astype(np.float64) @ _bf16_round(b).astype(np.float64), rtol=1e-5, atol=1e-5)
    # a problem the bf16 kernel cannot take (tiny) runs exactly, on the fp32 data
    with light.no_grad():
        fake_device.mode_log = []
        small = (CudaTensor.from_numpy(a[:8, :8]) @ CudaTensor.from_numpy(b[:8, :8])).numpy()
        assert fake_device.mode_log == [rt.GEMM_FP32_SIMT]
    np.testing.assert_allclose(small, a[:8, :8] @ b[:8, :8], rtol=1e-5, atol=1e-6)


def test_optimizer_step_drops_parameter_staging_copy_host_logic(fake_device, bf16_mode):
    import lightgrad_b200.nn as nn
    with nn.use_tensor(CudaTensor):
        np.random.seed(0)
        lin = nn.Linear(64, 256)
    opt = light.optim.SGD(lin.parameters(), lr=0.1)
    x = CudaTensor.from_numpy(np.random.RandomState(2).uniform(-1, 1, (256, 64)).astype(np.float32))
    y1 = lin(x)
    assert opt.arena.param_buf._bf16 is not None
    opt.zero_grad()
    y1.sum().backward()
    opt.step()
    assert opt.arena.param_buf._bf16 is None
    y2 = lin(x)
    w = lin.weight.numpy()
    want = _bf16_round(x.numpy()).astype(np.float64) @ _bf16_round(w).astype(np.float64).T + lin.bias.numpy()
    np.testing.assert_allclose(y2.numpy(), want, rtol=1e-4, atol=1e-4)
