"""CudaTensor -- the sm_100a device tensor behind lightgrad's AbstractTensor plugin API.

Plays the role OpenCLTensor plays in the reference (lightgrad/autograd/opencl/tensor.py:18-116):
a strided view (shape, strides, offset, dtype) over a pooled device buffer.  Because this module is
``<pkg>.cuda.tensor`` every tensor class gains a ``.cuda()`` converter (tensor.py:11-14 of the
reference).  All arithmetic is done by the kernels in ``lightgrad_b200/csrc`` through the ctypes
C-ABI; there is no numpy/torch compute path -- numpy is only the host-side container for
``from_numpy`` / ``numpy``.
"""
import ctypes as C
import numpy as np
from ..tensor import AbstractTensor
from . import runtime as rt

_F32 = np.dtype(np.float32)


def contiguous_strides(shape):
    st, acc = [], 1
    for s in reversed(shape):
        st.append(acc)
        acc *= s
    return tuple(reversed(st))


def _is_contiguous(shape, strides):
    acc = 1
    for s, st in zip(reversed(shape), reversed(strides)):
        if s != 1 and st != acc:
            return False
        acc *= s
    return True


def _prod(shape):
    n = 1
    for s in shape:
        n *= s
    return n


def i64arr(values):
    return (C.c_int64 * len(values))(*values)


class CudaTensor(AbstractTensor):
    # softmax(axis, scale=...) multiplies its input first (lets attention fuse the 1/sqrt(d))
    has_scaled_softmax = True
    # dot(other, out_layout='rhs') writes the product in the memory order of ``other``
    dot_supports_layout = True

    def __init__(self, data, shape=None, strides=None, offset=0, dtype=np.float32, requires_grad=True):
        if not isinstance(data, (rt.Buffer, rt.ArenaSlice)):
            # convenience: CudaTensor(ndarray | list | CudaTensor [, dtype=...]) uploads, like CpuTensor(data)
            if isinstance(data, CudaTensor):
                data = data.numpy()
            arr = np.ascontiguousarray(np.asarray(data, dtype=dtype))
            src = CudaTensor.from_numpy(arr, requires_grad=requires_grad)
            data, shape, strides, offset, dtype = src.data, src.shape, src.strides, 0, arr.dtype
        AbstractTensor.__init__(self, data=data, requires_grad=requires_grad)
        self._dtype = dtype if isinstance(dtype, np.dtype) else np.dtype(dtype)
        self._code = rt.DTYPE_CODE[self._dtype]
        self._shape = tuple(int(s) for s in shape)
        self._strides = tuple(strides) if strides is not None else contiguous_strides(self._shape)
        self._offset = offset
        self._numel = _prod(self._shape)
        self._contig = _is_contiguous(self._shape, self._strides)
        assert len(self._shape) == len(self._strides)
        assert len(self._shape) <= rt.MAX_DIMS, "CudaTensor supports at most %d dims" % rt.MAX_DIMS

    # -- metadata ------------------------------------------------------------
    @property
    def dtype(self):
        return self._dtype

    @property
    def shape(self):
        return self._shape

    @property
    def strides(self):
        return self._strides

    @property
    def offset(self):
        return self._offset

    @property
    def ptr(self):
        """Device address of element 0 of this view."""
        return self._data.ptr + self._offset * self._dtype.itemsize

    def numel(self):
        return self._numel

    def is_contiguous(self):
        return self._contig

    def __len__(self):
        return self._shape[0]

    def __repr__(self):
        return "CudaTensor(shape=%s, dtype=%s%s)" % (self._shape, self._dtype.name,
                                                     "" if self._contig else ", strided")

    # -- construction ----------------------------------------------------------
    @classmethod
    def _new(cls, shape, dtype=_F32, requires_grad=True):
        """Uninitialised contiguous tensor that solely owns its buffer (adoptable as a gradient)."""
        dtype = dtype if isinstance(dtype, np.dtype) else np.dtype(dtype)
        shape = tuple(shape)
        buf = rt.Buffer(max(_prod(shape), 1) * dtype.itemsize)
        t = cls(buf, shape, None, 0, dtype, requires_grad)
        t._temp = True
        return t

    @classmethod
    def empty(cls, shape, dtype=np.float32, requires_grad=True):
        shape = (shape,) if isinstance(shape, (int, np.integer)) else tuple(shape)
        t = cls._new(shape, dtype, requires_grad)
        t._temp = False
        return t

    @classmethod
    def zeros(cls, shape, dtype=np.float32, requires_grad=True):
        t = cls.empty(shape, dtype, requires_grad)
        rt.api.memset(t._data.ptr, 0, t._numel * t._dtype.itemsize)
        return t

    @classmethod
    def ones(cls, shape, dtype=np.float32, requires_grad=True):
        t = cls.empty(shape, dtype, requires_grad)
        t._fill_value(1)
        return t

    @classmethod
    def uniform(cls, low, high, shape, dtype=np.float32, requires_grad=True):
        # host RNG (numpy's global generator, as the reference's CpuTensor.uniform) then one upload,
        # so a seeded run initialises identically on every backend
        shape = (shape,) if isinstance(shape, (int, np.integer)) else tuple(shape)
        return cls.from_numpy(np.random.uniform(low, high, size=shape).astype(dtype), requires_grad=requires_grad)

    @classmethod
    def from_numpy(cls, a, requires_grad=True):
        a = np.asarray(a)
        if a.dtype == np.bool_:
            a = a.astype(np.uint8)
        if a.dtype not in rt.DTYPE_CODE:
            raise TypeError("CudaTensor does not support dtype %s" % a.dtype)
        src = np.ascontiguousarray(a)
        t = cls.empty(a.shape, a.dtype, requires_grad)
        if src.nbytes:
            rt.api.memcpy_h2d(t._data.ptr, src.ctypes.data, src.nbytes)
        return t

    def numpy(self):
        src = self if self._contig else self.contiguous()
        out = np.empty(self._shape, dtype=self._dtype)
        rt.api.memcpy_d2h(out.ctypes.data, src.ptr, out.nbytes)
        return out

    def copy(self, requires_grad=True):
        out = CudaTensor._new(self._shape, self._dtype, requires_grad)
        if self._numel:
            if self._contig:
                rt.api.memcpy_d2d(out._data.ptr, self.ptr, self._numel * self._dtype.itemsize)
            else:
                rt.api.cast(self._code, self._code, len(self._shape), i64arr(self._shape), self.ptr,
                            i64arr(self._strides), out._data.ptr, None)
        return out

    def contiguous(self):
        """Self if already dense, else a dense copy (one strided-gather kernel)."""
        return self if self._contig else self.copy(self._requires_grad)

    def astype(self, dtype):
        dtype = np.dtype(dtype)
        if dtype == self._dtype:
            return self
        out = CudaTensor._new(self._shape, dtype, self._requires_grad)
        if self._numel:
            rt.api.cast(self._code, out._code, len(self._shape), i64arr(self._shape), self.ptr,
                        i64arr(self._strides), out._data.ptr, None)
        return out

    # -- views ---------------------------------------------------------------
    def _view(self, shape, strides, offset=None):
        v = CudaTensor(self._data, shape, strides, self._offset if offset is None else offset,
                       self._dtype, True)
        return v

    def _fill_value(self, val):
        if self._numel == 0:
            return
        if val == 0 and self._contig:
            rt.api.memset(self.ptr, 0, self._numel * self._dtype.itemsize)
        elif self._code in (rt.F32, rt.F64):
            if self._contig:
                rt.api.ew_flat(rt.EW['FILL'], self._code, self.ptr, None, None, self.ptr, self._numel, float(val))
            else:
                shp, st = i64arr(self._shape), i64arr(self._strides)
                rt.api.ew(rt.EW['FILL'], self._code, len(self._shape), shp, self.ptr, st, None, None, None, None,
                          self.ptr, st, float(val))
        else:
            # integer tensors: stage the constant on the host (never on a hot path)
            host = np.full(self._shape, val, dtype=self._dtype)
            tmp = CudaTensor.from_numpy(host)
            rt.api.cast(self._code, self._code, len(self._shape), i64arr(self._shape), tmp.ptr, None,
                        self.ptr, i64arr(self._strides))


from . import ops  # noqa: E402,F401  (registers every operator on CudaTensor)
