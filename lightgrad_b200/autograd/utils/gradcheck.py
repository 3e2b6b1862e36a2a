"""Gradient checking: analytic Jacobians from the autograd graph against central differences.

Public contract (the four names the reference exports from lightgrad/autograd/utils/gradcheck.py and
that test/common.py:129 of the reference calls): ``jacobian``, ``numerical_jacobian``, ``gradcheck``,
``assert_gradcheck``; Jacobians are ``(x.numel(), y.numel())`` numpy arrays in C order of the logical
shapes.  Written for a device backend, where every host round trip costs a synchronisation:

* the analytic Jacobian runs ``f`` ONCE and pulls out one column per backward pass by seeding the
  pass with a one-hot cotangent, expressed as ``sum(y * e_j)`` with ``e_j`` a constant -- no
  ``__getitem__`` node per output element;
* the numerical Jacobian perturbs a host copy of ``x`` (one upload per evaluation, no per-element
  ``__setitem__`` launch on a device tensor) and differences the two results in float64.
"""
import numpy as np
from ..tensor import AbstractTensor
from ..grads import Gradients


def _host(t):
    return np.asarray(t.numpy())


def _one_hot_like(y_host, j):
    e = np.zeros(y_host.size, dtype=y_host.dtype if y_host.dtype.kind == 'f' else np.float32)
    e[j] = 1
    return e.reshape(y_host.shape)


def jacobian(f, x):
    """d f(x)[j] / d x[i] at entry (i, j), from ``x.grad`` after one backward pass per output element."""
    if not (isinstance(x, AbstractTensor) and x.requires_grad):
        raise AssertionError("jacobian: x must be a tensor that requires grad")
    y = f(x)
    if not (isinstance(y, AbstractTensor) and y.requires_grad):
        raise AssertionError("jacobian: f must return a tensor that requires grad")
    T = y.__class__
    y_host = _host(y)
    cols = []
    for j in range(y_host.size):
        seed = T.from_numpy(_one_hot_like(y_host, j), requires_grad=False)
        probe = (y * seed).sum()
        probe.zero_grad(traverse_graph=True)      # clears x, y and everything between them
        probe.backward()
        cols.append(_host(x.grad).reshape(-1).astype(x.dtype, copy=True))
    if not cols:
        return np.empty((x.numel(), 0), dtype=x.dtype)
    return np.stack(cols, axis=1)


def numerical_jacobian(f, x, eps=1e-4):
    """Central differences (f(x + eps e_i) - f(x - eps e_i)) / (2 eps), one row per input element."""
    if not isinstance(x, AbstractTensor):
        raise AssertionError("numerical_jacobian: x must be a tensor")
    T = x.__class__
    base = np.ascontiguousarray(_host(x))
    rows = []
    with Gradients.no_grad():
        for i in range(base.size):
            outs = []
            for sign in (1.0, -1.0):
                moved = base.copy()
                moved.reshape(-1)[i] += np.asarray(sign * eps, dtype=base.dtype)
                out = f(T.from_numpy(moved))
                if not isinstance(out, AbstractTensor):
                    raise AssertionError("numerical_jacobian: f must return a tensor")
                outs.append(_host(out).reshape(-1).astype(np.float64))
            rows.append((outs[0] - outs[1]) / (2.0 * eps))
    if not rows:
        return np.empty((0, f(x).numel()), dtype=x.dtype)
    return np.stack(rows, axis=0).astype(x.dtype)


def gradcheck(f, x, eps=1e-3, atol=5e-4, rtol=5e-4):
    return bool(np.allclose(jacobian(f, x), numerical_jacobian(f, x, eps), atol=atol, rtol=rtol))


def assert_gradcheck(f, x, eps=1e-3, atol=5e-4, rtol=5e-4):
    np.testing.assert_allclose(jacobian(f, x), numerical_jacobian(f, x, eps), atol=atol, rtol=rtol)
