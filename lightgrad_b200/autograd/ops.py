"""Backend-independent composite operators.

Everything here is expressed through the primitive set a backend must provide
({neg, add, mul, pow, exp, max, sum, fill, reshape, transpose, getitem,
setitem}), exactly as in the reference (lightgrad/autograd/ops.py:10-148).
Backends are free to shadow any of them with native kernels through
``register_op(..., overwrite=True)``; the cuda backend does so for sub, div,
sigmoid, tanh, softmax and mean.
"""
from .tensor import AbstractTensor
from .func import Function, WrapperFunction


# -- python operators -> named ops (ops.py:10-20) ----------------------------
def _bind(method, opname):
    setattr(AbstractTensor, method, lambda self, *a: getattr(self, opname)(*a))


for _m, _o in (('__neg__', 'neg'), ('__pow__', 'pow'),
               ('__add__', 'add'), ('__iadd__', 'add'), ('__radd__', 'add'),
               ('__mul__', 'mul'), ('__imul__', 'mul'), ('__rmul__', 'mul')):
    _bind(_m, _o)


def _composite(*names):
    """Register ``fn`` as a WrapperFunction under each of ``names`` (first name None -> fn.__name__)."""
    def deco(fn):
        op = WrapperFunction.from_function(fn)
        for n in names:
            AbstractTensor.register_op(n if n is not None else fn.__name__, op)
        return op
    return deco


# -- arithmetic built from add/neg and mul/pow (ops.py:22-47) -----------------
@_composite(None, '__sub__', '__isub__')
def sub(a, b):
    return a + (-b)


@_composite(None, '__truediv__', '__itruediv__')
def div(a, b):
    return a * (b ** -1)


@_composite('__rsub__')
def rsub(b, a):
    # a - b where only b is guaranteed to be a tensor
    return b.__class__.sub(a, b)


@_composite('__rtruediv__')
def rdiv(b, a):
    return b.__class__.div(a, b)


# -- activations (ops.py:52-66) ----------------------------------------------
@_composite(None)
def sigmoid(t):
    return 1 / (1 + t.neg().exp())


@_composite(None)
def tanh(t):
    return t.sigmoid() * 2 - 1


@_composite(None)
def softmax(t, axis=-1):
    e = (t - t.max(axis=axis, keepdims=True)).exp()
    return e / e.sum(axis=axis, keepdims=True)


# -- reductions (ops.py:71-75) -----------------------------------------------
@_composite(None)
def mean(t, axis=None, keepdims=False):
    s = t.sum(axis=axis, keepdims=keepdims)
    return s * (s.numel() / t.numel())


# -- padding and pooling windows (ops.py:79-148) -----------------------------
def _full(n):
    return slice(0, n)


@AbstractTensor.register_op()
class pad(Function):
    """Constant padding of the trailing ``len(dims)`` axes."""

    def forward(ctx, t, padding, dims=(-2, -1), value=0.0):
        lo, hi = padding if isinstance(padding, tuple) else (padding, padding)
        k = len(dims)
        ctx.save_for_backward(lo, hi, dims)
        head, tail = t.shape[:-k], t.shape[-k:]
        out = t.__class__.empty(head + tuple(lo + s + hi for s in tail), dtype=t.dtype)
        out = out.fill(value).detach()
        window = tuple(_full(s) for s in head) + tuple(slice(lo, lo + s) for s in tail)
        out[window] = t
        return out

    def backward(ctx, out_grad):
        lo, hi, dims = ctx.get_saved_tensors()
        window = [_full(s) for s in out_grad.shape]
        for d in dims:
            window[d] = slice(lo, out_grad.shape[d] - hi)
        return out_grad[tuple(window)]


@AbstractTensor.register_op()
class pool(Function):
    """Rearrange non-overlapping windows: (..., H, W) -> (kh*kw, ..., H/kh, W/kw)."""

    def forward(ctx, t, kernel=(2, 2)):
        k, r = len(kernel), len(t.shape)
        head = t.shape[:-k]
        fit = head + tuple((d // w) * w for d, w in zip(t.shape[-k:], kernel))
        x = t[tuple(_full(d) for d in fit)]
        ctx.save_for_backward(kernel, fit, t.shape)
        split = ()
        for d, w in zip(fit[-k:], kernel):
            split += (d // w, w)
        x = x.reshape(*head, *split)
        # window-element axes first, then batch axes, then window-position axes
        order = tuple(range(r - k + 1, r + k, 2)) + tuple(range(r - k)) + tuple(range(r - k, r + k, 2))
        x = x.transpose(*order)
        cells = 1
        for w in kernel:
            cells *= w
        return x.reshape(cells, *head, *(d // w for d, w in zip(fit[-k:], kernel)))

    def backward(ctx, out_grad):
        kernel, fit, in_shape = ctx.get_saved_tensors()
        k, r = len(kernel), len(fit)
        order = tuple(range(k, r))
        for i in range(k):
            order += (r + i, i)
        g = out_grad.reshape(*kernel, *out_grad.shape[1:])
        g = g.transpose(*order).reshape(*fit)
        if fit != in_shape:
            full = out_grad.__class__.zeros(in_shape)
            full[tuple(_full(d) for d in fit)] = g
            return full
        return g


@_composite(None)
def max_pool(t, kernel=(2, 2)):
    return t.pool(kernel=kernel).max(axis=0, keepdims=False)


@_composite(None)
def min_pool(t, kernel=(2, 2)):
    return t.pool(kernel=kernel).min(axis=0, keepdims=False)


@_composite(None)
def mean_pool(t, kernel=(2, 2)):
    return t.pool(kernel=kernel).mean(axis=0, keepdims=False)
