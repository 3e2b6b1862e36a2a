"""Test helpers shaped like the reference's test/common.py (:5-128): the same three oracles
(numpy, the CPU tensor, finite differences) and the same input generator, including its quirks --
`broadcast=True` re-draws each operand with one dim collapsed to 1 as FLOAT64 (common.py:25) and
`transpose=True` feeds fully reversed-axes views (common.py:33-38)."""
import numpy as np
from lightgrad_b200.autograd.tensor import AbstractTensor
from lightgrad_b200.autograd.utils.gradcheck import assert_gradcheck


def yield_input_pairs(cls, shapes, lowhigh=(-1, 1), dtype=np.float32, broadcast=False, transpose=False):
    assert len(lowhigh) == 2 and issubclass(cls, AbstractTensor)
    np_arrays = [np.random.uniform(*lowhigh, size=shape).astype(dtype) for shape in shapes]
    cls_arrays = [cls.from_numpy(arr) for arr in np_arrays]
    yield np_arrays, cls_arrays
    if broadcast:
        for i, shape in enumerate(shapes):
            for j in range(len(shape)):
                collapsed = shape[:j] + (1,) + shape[j + 1:]
                arr = np.random.uniform(*lowhigh, size=collapsed)      # float64 on purpose
                yield (np_arrays[:i] + [arr] + np_arrays[i + 1:],
                       cls_arrays[:i] + [cls.from_numpy(arr)] + cls_arrays[i + 1:])
    if transpose:
        for i, (arr, t, shape) in enumerate(zip(np_arrays, cls_arrays, shapes)):
            perm = list(reversed(range(len(shape))))
            yield (np_arrays[:i] + [arr.transpose(*perm)] + np_arrays[i + 1:],
                   cls_arrays[:i] + [t.transpose(*perm)] + cls_arrays[i + 1:])


def _resolve(owner_a, owner_b, fn_or_name):
    if isinstance(fn_or_name, str):
        return getattr(owner_a, fn_or_name), getattr(owner_b, fn_or_name)
    return fn_or_name, fn_or_name


def compare_with_numpy(cls, fn_or_name, shapes, lowhigh=(-1, 1), dtype=np.float32, broadcast=False,
                       transpose=False, rtol=1e-5, atol=1e-5, **kwargs):
    np_fn, cls_fn = _resolve(np, cls, fn_or_name)
    for np_arrays, cls_arrays in yield_input_pairs(cls, shapes, lowhigh, dtype, broadcast, transpose):
        want = np_fn(*np_arrays, **kwargs)
        got = cls_fn(*cls_arrays, **kwargs).numpy()
        assert got.shape == np.shape(want)
        np.testing.assert_allclose(want, got, rtol=rtol, atol=atol)


def compare_with_cpu(cls, fn_or_name, shapes, lowhigh=(-1, 1), dtype=np.float32, broadcast=False,
                     transpose=False, rtol=1e-3, atol=1e-3, **kwargs):
    from oracle import CpuTensor
    cpu_fn, cls_fn = _resolve(CpuTensor, cls, fn_or_name)
    for np_arrays, cls_arrays in yield_input_pairs(cls, shapes, lowhigh, dtype, broadcast, transpose):
        cpu_arrays = [CpuTensor.from_numpy(arr) for arr in np_arrays]
        want = cpu_fn(*cpu_arrays, **kwargs).numpy()
        got = cls_fn(*cls_arrays, **kwargs).numpy()
        np.testing.assert_allclose(want, got, rtol=rtol, atol=atol)


def check_gradients(cls, fn_or_name, shapes, lowhigh=(-1, 1), dtype=np.float32, broadcast=False, transpose=False,
                    eps=1e-3, tol=5e-4, **kwargs):
    fn = getattr(cls, fn_or_name) if isinstance(fn_or_name, str) else fn_or_name
    for _, cls_arrays in yield_input_pairs(cls, shapes, lowhigh, dtype, broadcast, transpose):
        for i, arr in enumerate(cls_arrays):
            f = lambda x: fn(*cls_arrays[:i], x, *cls_arrays[i + 1:], **kwargs)  # noqa: E731
            assert_gradcheck(f=f, x=arr, eps=eps, atol=tol, rtol=tol)
