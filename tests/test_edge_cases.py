"""Edge cases of the tensor API the reference's suites touch implicitly: empty and zero-dim tensors, ragged
shapes that defeat the vectorised paths, mixed dtypes (test/common.py:25 feeds float64 operands), every
integer index dtype, negative and duplicate indices, strided views as operands, in-place ops on views."""
import numpy as np
import pytest
import lightgrad_b200 as light
from lightgrad_b200 import CudaTensor as T


@pytest.fixture(params=["fake", pytest.param("gpu", marks=pytest.mark.gpu)])
def device(request):
    np.random.seed(7)
    request.getfixturevalue("fake_device" if request.param == "fake" else "cuda")
    return request.param


def test_empty_and_zero_dim(device):
    e = T.zeros((0, 5))
    assert (e + 1.0).shape == (0, 5) and e.exp().numpy().shape == (0, 5)
    assert (e @ T.ones((5, 3))).shape == (0, 3)
    s = T.from_numpy(np.float32(3.0))
    assert s.shape == () and s.numel() == 1 and (s * 2.0).item() == 6.0
    v = T.from_numpy(np.arange(6, dtype=np.float32))
    assert v[3].shape == () and v[3].item() == 3.0 and v.sum().shape == ()
    assert v.max(keepdims=True).shape == (1,)
    k0 = T.ones((3, 0)) @ T.ones((0, 4))                # K == 0 -> zeros
    np.testing.assert_array_equal(k0.numpy(), np.zeros((3, 4), dtype=np.float32))


@pytest.mark.parametrize("shape", [(1,), (3,), (5, 7), (129, 3), (2, 3, 5, 7), (1, 1, 1), (1023,), (33, 31)])
def test_ragged_shapes_elementwise_and_reduce(device, shape):
    a = np.random.uniform(0.5, 2, shape).astype(np.float32)
    b = np.random.uniform(0.5, 2, shape).astype(np.float32)
    A, B = T.from_numpy(a), T.from_numpy(b)
    np.testing.assert_allclose(((A * B + A) / B - A.log()).numpy(), (a * b + a) / b - np.log(a), rtol=2e-6)
    for axis in [None] + list(range(len(shape))):
        np.testing.assert_allclose(A.sum(axis=axis).numpy(), a.sum(axis=axis), rtol=2e-6)
        np.testing.assert_array_equal(A.max(axis=axis).numpy(), a.max(axis=axis))
    # the same through a sliced (offset, unaligned) view
    if shape[-1] > 2:
        np.testing.assert_allclose((A[..., 1:] + B[..., :-1]).numpy(), a[..., 1:] + b[..., :-1], rtol=1e-6)


def test_mixed_dtypes_promote_like_numpy(device):
    a32 = np.random.uniform(-1, 1, (4, 5)).astype(np.float32)
    a64 = np.random.uniform(-1, 1, (1, 5))                                   # float64, as test/common.py:25
    out = T.from_numpy(a32) * T.from_numpy(a64)
    assert out.dtype == np.float64
    np.testing.assert_allclose(out.numpy(), a32 * a64, rtol=1e-12)
    assert (T.from_numpy(a32) + 1).dtype == np.float32                        # python scalars are weak
    i = T.from_numpy(np.arange(5, dtype=np.int32))
    assert (i * 0.5).dtype == np.float32
    np.testing.assert_allclose((T.from_numpy(a64) @ T.from_numpy(a64.T.copy())).numpy(), a64 @ a64.T, rtol=1e-12)


@pytest.mark.parametrize("idt", [np.int16, np.int32, np.int64, np.uint8, np.int8])
def test_index_dtypes_negative_and_duplicates(device, idt):
    x = np.random.uniform(-1, 1, (20, 6)).astype(np.float32)
    idx = np.array([3, 0, 19, 3, 7], dtype=idt)
    X = T.from_numpy(x)
    np.testing.assert_array_equal(X[T.from_numpy(idx, requires_grad=False)].numpy(), x[idx])
    if np.issubdtype(idt, np.signedinteger):
        neg = np.array([-1, -20, 4], dtype=idt)
        np.testing.assert_array_equal(X[T.from_numpy(neg, requires_grad=False)].numpy(), x[neg])
    X.zero_grad()
    X[idx].sum().backward()
    want = np.zeros_like(x)
    np.add.at(want, idx.astype(np.int64), 1.0)
    np.testing.assert_array_equal(X.grad.numpy(), want)                        # duplicates accumulate (scatter-add)


def test_views_as_operands_and_inplace_on_views(device):
    x = np.random.uniform(-1, 1, (6, 8)).astype(np.float32)
    X = T.from_numpy(x)
    xt = X.transpose(1, 0)
    np.testing.assert_allclose((xt @ X).numpy(), x.T @ x, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(xt.relu().sum(axis=1).numpy(), np.maximum(x.T, 0).sum(axis=1), rtol=1e-6)
    with light.no_grad():
        v = X[1:5, ::2]
        v += 1.0                                        # writes through to X's storage
        v *= T.from_numpy(np.full((4, 1), 2.0, dtype=np.float32))
    x[1:5, ::2] = (x[1:5, ::2] + 1.0) * 2.0
    np.testing.assert_allclose(X.numpy(), x, rtol=1e-6)
    c = xt.contiguous()
    assert c.is_contiguous() and not xt.is_contiguous()
    np.testing.assert_array_equal(c.numpy(), x.T)
    np.testing.assert_array_equal(X.reshape(2, 3, 8).transpose(1, 0, 2).reshape(6, 8).numpy(),
                                  x.reshape(2, 3, 8).transpose(1, 0, 2).reshape(6, 8))


def test_errors_are_python_exceptions(device):
    a, b = T.ones((3, 4)), T.ones((5, 4))
    with pytest.raises(ValueError):
        a + b
    with pytest.raises(ValueError):
        a @ b
    with pytest.raises(IndexError):
        a[7]
    with pytest.raises(RuntimeError):
        a.backward()                                     # no graph -> silently returns; with a graph: item tensors only
        (a * 2).backward()
