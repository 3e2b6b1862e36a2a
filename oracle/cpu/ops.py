"""ORACLE (test infrastructure): binds the numpy rules of ``oracle/np_ops.py`` to CpuTensor.

Follows the adapter of the reference (cpu/ops.py:8-21): operands are unwrapped to
ndarrays, the result is wrapped in a fresh CpuTensor of the result's dtype
(always ``requires_grad=True``, cpu/ops.py:14).
"""
import numpy as np
from lightgrad_b200.autograd.func import Function
from .tensor import CpuTensor
from ..np_ops import RULES


def _raw(v):
    if isinstance(v, CpuTensor):
        return v.data
    if isinstance(v, tuple):
        return tuple(_raw(x) for x in v)
    return v


def _make(name, r):
    def forward(ctx, *args, **kwargs):
        out, saved = r.fwd(*[_raw(a) for a in args], **{k: _raw(v) for k, v in kwargs.items()})
        ctx.save_for_backward(saved)
        if r.inplace:
            # same storage, new wrapper object (cpu/ops.py:13-14 wraps whatever forward returned)
            return CpuTensor(args[0].data, dtype=args[0].data.dtype)
        out = np.asarray(out)
        return CpuTensor(out, dtype=out.dtype)
    body = {'forward': forward}
    if r.bwd is not None:
        def backward(ctx, out_grad):
            saved, = ctx.get_saved_tensors()
            grads = r.bwd(saved, out_grad.data)
            return tuple(CpuTensor(np.asarray(g), dtype=np.asarray(g).dtype) for g in grads)
        body['backward'] = backward
    return type(name, (Function,), body)


_ALIASES = {'transpose': ('transpose', 'T'), 'dot': ('dot', '__matmul__')}
for _name, _rule in RULES.items():
    _op = _make(_name.strip('_'), _rule)
    for _n in _ALIASES.get(_name, (_name,)):
        CpuTensor.register_op(_n, _op, overwrite=True)
