"""Modules: parameter registry plus the layers used by examples/mnist.py and examples/bert.py.

API of the reference's lightgrad/nn.py (Module :4-76, ModuleList :78-88, Linear :90-96,
Conv2d :98-107, LayerNorm :109-124) plus ``Embedding`` (examples/bert.py:14-21 of the reference,
kept on the device here).  Layers dispatch to a backend's fused operator when the tensor class
provides one (``linear``, ``layernorm``) and otherwise compose primitives exactly as the
reference does -- so the same module runs on any tensor backend registered with the core.
"""
import numpy as np
from contextlib import contextmanager
from . import autograd
from .autograd import AbstractTensor

_default_tensor = [None]


def default_tensor():
    """Tensor class new parameters are created with (``autograd.Tensor`` unless overridden)."""
    return _default_tensor[0] if _default_tensor[0] is not None else autograd.Tensor


@contextmanager
def use_tensor(cls):
    """Create the parameters of modules built inside the block with tensor class ``cls``."""
    prev = _default_tensor[0]
    _default_tensor[0] = cls
    try:
        yield cls
    finally:
        _default_tensor[0] = prev


class Module(object):

    def __init__(self):
        object.__setattr__(self, '_params', {})
        object.__setattr__(self, '_children', {})

    def forward(self, x):
        raise NotImplementedError()

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    def __setattr__(self, name, val):
        if isinstance(val, (AbstractTensor, Module)):
            self.register_param_or_module(name, val)
        object.__setattr__(self, name, val)

    def register_param_or_module(self, name, val):
        if isinstance(val, AbstractTensor):
            self._params[name] = val
        elif isinstance(val, Module):
            self._children[name] = val
        return val

    def unregister_param_or_module(self, name):
        if name in self._params:
            return self._params.pop(name)
        if name in self._children:
            return self._children.pop(name)

    def parameters(self):
        for p in self._params.values():
            yield p
        for m in self._children.values():
            for p in m.parameters():
                yield p

    def named_parameters(self, prefix="", separator="."):
        base = (prefix + separator) if len(prefix) > 0 else ""
        for name, p in self._params.items():
            yield base + name, p
        for name, m in self._children.items():
            for item in m.named_parameters(prefix=base + name, separator=separator):
                yield item

    def map_parameters(self, fn):
        for key in list(self._params):
            setattr(self, key, fn(self._params[key]))
        for m in self._children.values():
            m.map_parameters(fn)
        return self

    def load_parameters(self, param_dict, prefix="", separator='.'):
        param_dict = dict(param_dict)
        base = (prefix + separator) if len(prefix) > 0 else ""
        for key, p in list(self._params.items()):
            full = base + key
            assert full in param_dict, "%s not found in param dict!" % full
            new_p = param_dict[full]
            if not isinstance(new_p, p.__class__):
                new_p = new_p.numpy() if isinstance(new_p, AbstractTensor) else new_p
                assert isinstance(new_p, np.ndarray), "Unexpected parameter type %s!" % new_p.__class__.__name__
                new_p = p.__class__.from_numpy(new_p)
            assert p.shape == new_p.shape, "Shapes do not align! (%s != %s)" % (p.shape, new_p.shape)
            setattr(self, key, new_p)
        for key, m in self._children.items():
            m.load_parameters(param_dict, prefix=base + key, separator=separator)


class ModuleList(Module, list):

    def __init__(self, *elements):
        Module.__init__(self)
        list.__init__(self, elements)
        for i, e in enumerate(elements):
            self.register_param_or_module(str(i), e)

    def __setitem__(self, i, e):
        assert i < len(self)
        self.unregister_param_or_module(str(i))
        self.register_param_or_module(str(i), e)
        return list.__setitem__(self, i, e)


class Linear(Module):
    """y = x W^T + b with W stored (out_feats, in_feats)."""

    def __init__(self, in_feats, out_feats, bias=True):
        Module.__init__(self)
        T = default_tensor()
        self.weight = T.xavier((out_feats, in_feats))
        self.bias = T.xavier((out_feats,)) if bias else None

    def forward(self, x):
        if hasattr(x, 'linear'):
            return x.linear(self.weight, self.bias) if self.bias is not None else x.linear(self.weight)
        y = x @ self.weight.T(1, 0)
        return (y + self.bias) if self.bias is not None else y


class Conv2d(Module):

    def __init__(self, in_channels, out_channels, kernelsize=3, stride=1, pad=None, bias=True):
        Module.__init__(self)
        T = default_tensor()
        self.w = T.xavier((out_channels, in_channels, kernelsize, kernelsize))
        self.b = T.xavier((1, out_channels, 1, 1)) if bias else None
        self.s, self.p = stride, (kernelsize // 2) if pad is None else pad

    def forward(self, x):
        y = (x.pad(self.p) if self.p > 0 else x).conv(self.w, strides=self.s)
        return (y + self.b) if self.b is not None else y


class LayerNorm(Module):

    def __init__(self, shape, eps=1e-5):
        Module.__init__(self)
        T = default_tensor()
        self.shape = tuple(shape) if isinstance(shape, (tuple, list)) else (shape,)
        self.eps = eps
        self.weight = T.ones(self.shape)
        self.bias = T.zeros(self.shape)

    def forward(self, x, residual=None):
        """``residual``: normalise ``x + residual`` (backends with a fused add + layer norm do it in one pass)."""
        k = len(self.shape)
        assert x.shape[-k:] == self.shape, "Shape mismatch in layer norm! (%s <-> %s)" % (x.shape, self.shape)
        if residual is not None:
            if k == 1 and hasattr(x, 'add_layernorm') and x.shape == residual.shape and x.dtype == residual.dtype:
                return x.add_layernorm(residual, self.weight, self.bias, eps=self.eps)
            x = x + residual
        if k == 1 and hasattr(x, 'layernorm'):
            return x.layernorm(self.weight, self.bias, eps=self.eps)
        axes = tuple(range(len(x.shape) - k, len(x.shape)))
        D = x - x.mean(axis=axes, keepdims=True)
        V = (D * D).mean(axis=axes, keepdims=True)
        return D / (V + self.eps).pow(1 / 2) * self.weight + self.bias


class Embedding(Module):
    """Row lookup ``weight[ids]``; differentiable (gather forward, scatter-add backward)."""

    def __init__(self, embedding_dim, vocab_size):
        Module.__init__(self)
        self.d, self.n = embedding_dim, vocab_size
        self.weight = default_tensor().xavier((vocab_size, embedding_dim))

    def forward(self, ids):
        return self.weight[ids]
