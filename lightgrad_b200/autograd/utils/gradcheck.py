"""Finite-difference gradient checking (API of lightgrad/autograd/utils/gradcheck.py:5-63).

``jacobian`` back-propagates one output element at a time (gradcheck.py:18-24),
``numerical_jacobian`` uses central differences with a one-hot perturbation
written through ``__setitem__`` (gradcheck.py:39-48).
"""
import numpy as np
from ..tensor import AbstractTensor
from ..grads import Gradients


def jacobian(f, x):
    assert isinstance(x, AbstractTensor) and x.requires_grad
    y = f(x)
    assert isinstance(y, AbstractTensor) and y.requires_grad
    n_in, n_out = x.numel(), y.numel()
    y = y.reshape(-1)
    J = np.empty((n_in, n_out), dtype=x.dtype)
    for j in range(n_out):
        y.zero_grad(traverse_graph=True)
        y[j].backward()
        J[:, j] = x.grad.reshape(-1).numpy()
    return J


@Gradients.no_grad()
def numerical_jacobian(f, x, eps=1e-4):
    assert isinstance(x, AbstractTensor)
    y = f(x)
    assert isinstance(y, AbstractTensor)
    n_in, n_out = x.numel(), y.numel()
    NJ = np.empty((n_in, n_out), dtype=x.dtype)
    for i, idx in enumerate(np.ndindex(x.shape)):
        h = x.__class__.zeros(x.shape)
        h[idx] = eps
        hi = f(x + h).reshape(-1)
        lo = f(x - h).reshape(-1)
        NJ[i, :] = (hi - lo).numpy() / (2 * eps)
    return NJ


def gradcheck(f, x, eps=1e-3, atol=5e-4, rtol=5e-4):
    return np.allclose(jacobian(f, x), numerical_jacobian(f, x, eps), atol=atol, rtol=rtol)


def assert_gradcheck(f, x, eps=1e-3, atol=5e-4, rtol=5e-4):
    return np.testing.assert_allclose(jacobian(f, x), numerical_jacobian(f, x, eps), atol=atol, rtol=rtol)
