"""ORACLE (test infrastructure, never shipped or measured as the product).

Plain-numpy restatement of the reference CPU tensor's operator arithmetic:
every rule is ``fwd(*arrays, **kw) -> (out, saved)`` and ``bwd(saved, g) ->
tuple(grads)``, and cites the reference lines it follows
(``/root/reference/lightgrad/autograd/cpu/ops.py`` unless another file is
named).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
cpu_baseline / ``--impl reference`` legs may import this package.

Parity pinning: ``tests/golden/*.npz`` were produced by the REAL reference
(imported from /root/reference in the build container by
``tests/golden/make_golden.py``); ``tests/test_oracle.py`` checks this
restatement against them.

Three rules intentionally differ from the reference as shipped (SURVEY.md
F2/F4c), and the golden generator applies the same three patches to the
reference through its own ``register_op(..., overwrite=True)``:
  * ``sum`` has a backward (reference: cpu/ops.py:288-293 has none),
  * ``dot`` backward swaps the last two axes instead of ``.T`` (cpu/ops.py:114-116
    is only right for 2-D operands),
  * ``getitem`` backward is scatter-ADD (cpu/ops.py:242-246 assigns, losing
    the gradient of duplicate indices).
"""
import numpy as np
from math import ceil


class Rule(object):
    def __init__(self, fwd, bwd=None, inplace=False):
        self.fwd, self.bwd, self.inplace = fwd, bwd, inplace


RULES = {}


def rule(name, bwd=None, inplace=False):
    def deco(fwd):
        RULES[name] = Rule(fwd, bwd, inplace)
        return fwd
    return deco


# ---------------------------------------------------------------- layout ----
def _transpose_bwd(axes, g):                       # cpu/ops.py:31-36
    if len(axes) == 0:
        return (np.transpose(g),)
    inv = np.argsort(axes)
    return (np.transpose(g, inv),)


@rule('transpose', _transpose_bwd)                 # cpu/ops.py:25-30
def _transpose(a, *axes):
    return np.transpose(a, axes if len(axes) else None), axes


@rule('reshape', lambda shape, g: (g.reshape(shape),))   # cpu/ops.py:38-47
def _reshape(a, *shape):
    return a.reshape(shape), a.shape


# ------------------------------------------------------------ arithmetic ----
@rule('neg', lambda s, g: (-g,))                   # cpu/ops.py:52-58
def _neg(a):
    return -a, None


@rule('add', lambda s, g: (g, g))                  # cpu/ops.py:60-66
def _add(a, b):
    return a + b, None


@rule('sub', lambda s, g: (g, -g))                 # cpu/ops.py:68-74
def _sub(a, b):
    return a - b, None


@rule('mul', lambda s, g: (g * s[1], s[0] * g))    # cpu/ops.py:76-84
def _mul(a, b):
    return a * b, (a, b)


@rule('div', lambda s, g: (g / s[1], -s[0] / s[1] ** 2 * g))   # cpu/ops.py:86-94
def _div(a, b):
    return a / b, (a, b)


def _pow_bwd(s, g):                                # cpu/ops.py:103-105
    a, b, y = s
    with np.errstate(all='ignore'):
        return b * (a ** (b - 1)) * g, g * y * np.log(a)


@rule('pow', _pow_bwd)                             # cpu/ops.py:96-102
def _pow(a, b):
    y = a ** b
    return y, (a, b, y)


def _dot_bwd(s, g):
    # PATCH (F2): reference uses a.T / b.T (cpu/ops.py:114-116), only valid for 2-D
    a, b = s
    return g @ np.swapaxes(b, -1, -2), np.swapaxes(a, -1, -2) @ g


@rule('dot', _dot_bwd)                             # cpu/ops.py:107-113
def _dot(a, b):
    return a @ b, (a, b)


# --------------------------------------------------------------- in place ----
@rule('__iadd__', inplace=True)                    # cpu/ops.py:120-126
def _iadd(t, other):
    t += other
    return t, None


@rule('__isub__', inplace=True)                    # cpu/ops.py:128-134
def _isub(t, other):
    t -= other
    return t, None


@rule('__imul__', inplace=True)                    # cpu/ops.py:136-142
def _imul(t, other):
    t *= other
    return t, None


@rule('__itruediv__', inplace=True)                # cpu/ops.py:144-150
def _idiv(t, other):
    t /= other
    return t, None


@rule('fill', inplace=True)                        # cpu/ops.py:152-157
def _fill(t, val):
    t.fill(val)
    return t, None


# ------------------------------------------------------------------ unary ----
@rule('sin', lambda x, g: (np.cos(x) * g,))        # cpu/ops.py:162-170
def _sin(t):
    return np.sin(t), t


@rule('cos', lambda x, g: (-np.sin(x) * g,))       # cpu/ops.py:172-180
def _cos(t):
    return np.cos(t), t


@rule('exp', lambda y, g: (y * g,))                # cpu/ops.py:182-191
def _exp(t):
    y = np.exp(t)
    return y, y


@rule('log', lambda x, g: ((1 / x) * g,))          # cpu/ops.py:193-201
def _log(t):
    return np.log(t), t


@rule('sigmoid', lambda y, g: (y * (1 - y) * g,))  # cpu/ops.py:203-212
def _sigmoid(t):
    y = 1 / (1 + np.exp(-t))
    return y, y


@rule('tanh', lambda y, g: ((1 - y ** 2) * g,))    # cpu/ops.py:214-223
def _tanh(t):
    y = np.tanh(t)
    return y, y


@rule('relu', lambda x, g: (g * (x >= 0),))        # cpu/ops.py:225-233 (grad 1 at x == 0)
def _relu(t):
    return np.maximum(t, 0.0), t


# --------------------------------------------------------------- indexing ----
def _getitem_bwd(s, g):
    # PATCH (F4c): scatter-add; the reference assigns (cpu/ops.py:242-246)
    shape, idx = s
    out = np.zeros(shape, dtype=np.float32)
    np.add.at(out, idx, g)
    return (out,)


@rule('__getitem__', _getitem_bwd)                 # cpu/ops.py:234-241
def _getitem(a, idx):
    return a[idx], (a.shape, idx)


@rule('__setitem__', inplace=True)                 # cpu/ops.py:248-255
def _setitem(a, idx, val):
    a[idx] = val
    return a, None


# ------------------------------------------------------------- reductions ----
def _extreme(npfn):
    def fwd(x, axis=None, keepdims=False):         # cpu/ops.py:260-267 / 274-281
        axis = tuple(range(x.ndim)) if axis is None else axis
        val = npfn(x, axis=axis, keepdims=True)
        return (val if keepdims else np.squeeze(val, axis=axis)), (x, val, axis, keepdims)

    def bwd(s, g):                                 # cpu/ops.py:268-272 / 282-286 (all ties get g)
        x, val, axis, keepdims = s
        if not keepdims:
            g = np.expand_dims(g, axis=axis)
        return (g * (x == val),)
    return fwd, bwd


_f, _b = _extreme(np.max)
RULES['max'] = Rule(_f, _b)
_f, _b = _extreme(np.min)
RULES['min'] = Rule(_f, _b)


def _sum_bwd(s, g):
    # PATCH (F2): the reference CPU tensor has no sum backward (cpu/ops.py:288-293);
    # this is the broadcast the OpenCL backend implements (opencl/ops.py:344-368)
    shape, axis, keepdims = s
    if axis is None:
        return (np.broadcast_to(g, shape).copy(),)
    if not keepdims:
        g = np.expand_dims(g, axis=axis)
    return (np.broadcast_to(g, shape).copy(),)


@rule('sum', _sum_bwd)                             # cpu/ops.py:288-291
def _sum(t, axis=None, keepdims=False):
    return t.sum(axis=axis, keepdims=keepdims), (t.shape, axis, keepdims)


# ------------------------------------------------------------ convolution ----
def _windows(t, kshape, strides):                  # cpu/ops.py:301-306
    n = len(kshape)
    shape = t.shape[:-n] + tuple((d - k) // s + 1 for d, k, s in zip(t.shape[-n:], kshape, strides)) + kshape
    st = t.strides[:-n] + tuple(ts * ws for ts, ws in zip(t.strides[-n:], strides)) + t.strides[-n:]
    return np.lib.stride_tricks.as_strided(t, shape=shape, strides=st)


def _conv_bwd(s, g):                               # cpu/ops.py:327-356
    flat_x, flat_w, in_shape, w_shape, strides = s
    n = len(w_shape) - 1
    flat_g = np.moveaxis(g, -n, -1).reshape(-1, w_shape[0])
    flat_xg = flat_g @ flat_w
    w_grad = (flat_g.T @ flat_x).reshape(w_shape)
    x_grad = np.zeros(in_shape)
    xw = _windows(x_grad, w_shape[1:], strides)
    src = flat_xg.reshape(xw.shape)
    skip = tuple(slice(0, d) for d in xw.shape[:-n])
    blk = tuple(st if ks < d else d for st, ks, d in zip(strides, w_shape[1:], in_shape[-n:]))
    for pos in np.ndindex(tuple(ceil(ks / st) for ks, st in zip(w_shape[1:], blk))):
        sel = skip + tuple(slice(i * st, i * st + st) for i, st in zip(pos, blk))
        xw[sel] += src[sel]
    return x_grad, w_grad


@rule('conv', _conv_bwd)                           # cpu/ops.py:308-325
def _conv(t, kernel, strides=1):
    n, m = kernel.ndim - 1, t.ndim
    strides = ((strides,) * n) if isinstance(strides, int) else \
        ((1,) + tuple(strides) if len(strides) == n - 1 else tuple(strides))
    assert m >= n == len(strides)
    x = _windows(t, kernel.shape[1:], strides)
    flat_x = x.reshape(-1, int(np.prod(kernel.shape[1:])))
    flat_w = kernel.reshape(kernel.shape[0], -1)
    y = (flat_x @ flat_w.T).reshape(*x.shape[:-n], -1)
    y = y.swapaxes(-n - 1, -1).squeeze(-1)
    return y, (flat_x, flat_w, t.shape, kernel.shape, strides)
