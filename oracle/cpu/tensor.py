"""ORACLE (test infrastructure): numpy tensor that plugs into lightgrad_b200's autograd core.

Restates ``lightgrad/autograd/cpu/tensor.py:4-46`` of the reference: float32 by
default, casts on construction, ``copy()`` silently returns float32
(cpu/tensor.py:39-40), ``numpy()`` returns the live array.  Because this module
is ``oracle.cpu.tensor`` every tensor gains a ``.cpu()`` converter, as in the
reference.
"""
import numpy as np
from lightgrad_b200.autograd.tensor import AbstractTensor


class CpuTensor(AbstractTensor):

    def __init__(self, data, dtype=np.float32, requires_grad=True):
        if isinstance(data, CpuTensor):
            data = data.data
        if isinstance(data, np.ndarray):
            if data.dtype != dtype:
                data = data.astype(dtype)
        else:
            data = np.asarray(data, dtype=dtype)
        AbstractTensor.__init__(self, data=data, requires_grad=requires_grad)

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def shape(self):
        return self.data.shape

    @staticmethod
    def empty(shape, *args, **kwargs):
        return CpuTensor(np.empty(shape), *args, **kwargs)

    @staticmethod
    def zeros(shape, *args, **kwargs):
        return CpuTensor(np.zeros(shape), *args, **kwargs)

    @staticmethod
    def ones(shape, *args, **kwargs):
        return CpuTensor(np.ones(shape), *args, **kwargs)

    @staticmethod
    def uniform(low, high, shape, *args, **kwargs):
        return CpuTensor(np.random.uniform(low, high, size=shape), *args, **kwargs)

    def copy(self, requires_grad=True):
        return CpuTensor(self.data.copy(), requires_grad=requires_grad)

    def numpy(self):
        return self.data

    @staticmethod
    def from_numpy(a, requires_grad=True):
        return CpuTensor(data=a, dtype=a.dtype, requires_grad=requires_grad)


from . import ops  # noqa: E402,F401
