"""Operators of the CUDA backend, registered on CudaTensor.

Same op names, argument meaning and gradient conventions as the reference's CPU backend
(lightgrad/autograd/cpu/ops.py) and the union with its OpenCL backend (opencl/ops.py); every
Function below cites the reference op it stands in for.  Conventions kept on purpose
(SURVEY.md F4): every result has requires_grad=True, relu' is 1 at 0, max/min send the full
gradient to every tie.  Divergence kept on purpose: getitem backward is scatter-ADD.

Host side only: shape/stride bookkeeping and kernel selection.  All arithmetic happens in
liblightgrad_b200.so.
"""
import os
import ctypes as C
import numpy as np
from ..func import Function
from ..tensor import AbstractTensor
from .tensor import CudaTensor, i64arr, contiguous_strides, _prod
from . import runtime as rt
from ..grads import Gradients

EW, RED = rt.EW, rt.RED
_SCALARS = (int, float, np.integer, np.floating, bool, np.bool_)

# ---------------------------------------------------------------------------------------------------
# matmul arithmetic mode: 'fp32' (exact FFMA, default), 'tf32' or 'bf16' (tcgen05 tensor cores)
_MODES = {'fp32': rt.GEMM_FP32_SIMT, 'tf32': rt.GEMM_TF32_TC, 'bf16': rt.GEMM_BF16_TC}
_matmul_mode = _MODES[os.environ.get('LG_MATMUL_MODE', 'fp32').lower()]


Gradients.after_backward.append(rt.side_join)

# A/B switch for measurements: LG_DISABLE=add_ln,mlp,attn,attn_epi,dx_acc turns the named fusions off
_DISABLED = set(filter(None, os.environ.get('LG_DISABLE', '').split(',')))


def set_matmul_mode(name):
    """Select how float32 matmuls are computed: 'fp32' | 'tf32' | 'bf16'.  Returns the previous mode name."""
    global _matmul_mode
    prev = get_matmul_mode()
    _matmul_mode = _MODES[name.lower()]
    return prev


def get_matmul_mode():
    return [k for k, v in _MODES.items() if v == _matmul_mode][0]


# ---------------------------------------------------------------------------------------------------
# bf16 tensor-core mode: operands are bf16 STAGING COPIES of the fp32 tensors (the tensors themselves, the
# accumulation and every result stay fp32).  A copy covers a whole device block (the tensor's buffer, or the
# optimizer arena a parameter lives in) and is shared by every view of it -- x, x^T, per-head slices, W and W^T
# all address the same copy with the same element strides -- so each block is converted once per step however many
# GEMMs read it (x: forward and dW; dY: dX and dW; W: forward and dX).  The copy is dropped whenever the block is
# written through this module (in-place operators, ``out=`` results, optimizer updates).
import weakref  # noqa: E402
_staged = weakref.WeakSet()


def _root(t):
    d = t._data
    return d.parent if isinstance(d, rt.ArenaSlice) else d


def _bf16_ptr(t):
    """Device address of view ``t`` inside the bf16 staging copy of its block (made on first use)."""
    root = _root(t)
    sh = root._bf16
    if sh is None:
        n = root.nbytes // 4
        sh = rt.Buffer(max(n, 8) * 2)
        rt.api.cast(rt.F32, rt.BF16, 1, i64arr((n,)), root.ptr, None, sh.ptr, None)
        root._bf16 = sh
        _staged.add(root)
    return sh.ptr + (t.ptr - root.ptr) // 2


def _written(t):
    """``t``'s block is about to be modified in place: its staging copy and any column sums a producer left for a
    matrix in it are stale."""
    root = _root(t)
    if root._bf16 is not None:
        root._bf16 = None
    if root._colsum is not None:
        root._colsum = None


def drop_staging_copies():
    """Forget every bf16 staging copy (before / after a CUDA-graph capture: a copy made outside the capture would
    not be refreshed by replays, one made inside it lives in the graph's private pool)."""
    for root in list(_staged):
        root._bf16 = None
    _staged.clear()


def _stage_for_side_stream(tensors):
    if _matmul_mode == rt.GEMM_BF16_TC:
        for t in tensors:
            if isinstance(t, CudaTensor) and t._code == rt.F32 and t._numel:
                _bf16_ptr(t)


rt.side_prepare = _stage_for_side_stream


# In bf16 mode every matmul belongs to a class -- F forward product, X gradient w.r.t. the activation (the chain the
# error travels along), W gradient w.r.t. a weight (ends in the optimizer), A the small batched attention products --
# and only the classes named in LG_BF16_CLASSES multiply bf16 operands; the others run in tf32 on the fp32 data.
# Default: attention stays tf32 (2 % of the flops; softmax probabilities keep 10 mantissa bits).
_BF16_CLASSES = set(os.environ.get('LG_BF16_CLASSES', 'FXW').upper())


def set_bf16_classes(classes):
    """Which matmul classes ('F', 'X', 'W', 'A') multiply bf16 operands in bf16 mode.  Returns the previous set."""
    global _BF16_CLASSES
    prev = ''.join(sorted(_BF16_CLASSES))
    _BF16_CLASSES = set(classes.upper())
    return prev


def _mode_for(cls, mode=None):
    m = _matmul_mode if mode is None else mode
    if m == rt.GEMM_BF16_TC and cls not in _BF16_CLASSES:
        return rt.GEMM_TF32_TC
    return m


def _launch_gemm(mode, code, d, a, b, out, bias, accumulate):
    """lg_gemm in the selected mode; in bf16 mode through the staging copies, or -- when the tensor-core kernel
    cannot take the problem (tiny, or strides that are not 16-byte multiples of bf16) -- exactly, on the fp32 data."""
    bias_ptr = bias.ptr if bias is not None else None
    if mode == rt.GEMM_BF16_TC:
        if code == rt.F32 and not (accumulate and bias is not None) and \
                rt.api.gemm_tc_supported(mode, rt.BF16, C.byref(d)):
            pa, pb = _bf16_ptr(a), _bf16_ptr(b)
            if not ((pa | pb | out.ptr) & 15):
                rt.api.gemm(mode, rt.BF16, C.byref(d), pa, pb, out.ptr, bias_ptr, 1 if accumulate else 0)
                return
        mode = rt.GEMM_FP32_SIMT
    rt.api.gemm(mode, code, C.byref(d), a.ptr, b.ptr, out.ptr, bias_ptr, 1 if accumulate else 0)


def matmul_mode_available(name):
    """True when the library can run a BERT-sized projection (4096 x 768 x 768, x @ W^T) in that mode on the
    tensor cores (always True for the exact 'fp32' mode)."""
    code = _MODES[name.lower()]
    if code == rt.GEMM_FP32_SIMT:
        return True
    d = rt.GemmDesc(4096, 768, 768, 1, 1, 0, 0, 768, 1, 0, 0, 1, 768, 0, 0, 768, 1)
    return bool(rt.ensure_device().gemm_tc_supported(code, rt.BF16 if code == rt.GEMM_BF16_TC else rt.F32, C.byref(d)))


# ---------------------------------------------------------------------------------------------------
# helpers
def _is_scalar(v):
    return isinstance(v, _SCALARS) or (isinstance(v, np.ndarray) and v.ndim == 0)


def _as_tensor(v, like):
    """Promote array-likes that reach an op (ndarray, list) to a CudaTensor."""
    if isinstance(v, CudaTensor):
        return v
    if isinstance(v, AbstractTensor):
        raise AssertionError("All Tensors must be of the same type!")
    a = np.asarray(v)
    if a.dtype == np.float64 and not isinstance(v, np.ndarray):
        a = a.astype(like.dtype)
    return CudaTensor.from_numpy(a, requires_grad=False)


def _float_like(t):
    """Arithmetic kernels are float32/float64; integer operands are promoted like numpy's true division."""
    if t._code in (rt.F32, rt.F64):
        return t
    return t.astype(np.float64 if t._code == rt.I64 else np.float32)


def _promote(a, b):
    a, b = _float_like(a), _float_like(b)
    if a._code != b._code:
        a, b = a.astype(np.float64), b.astype(np.float64)
    return a, b


def _bshape(sa, sb):
    if sa == sb:
        return sa
    n = max(len(sa), len(sb))
    pa, pb = (1,) * (n - len(sa)) + sa, (1,) * (n - len(sb)) + sb
    out = []
    for x, y in zip(pa, pb):
        if x != y and x != 1 and y != 1:
            raise ValueError("operands could not be broadcast together with shapes %s %s" % (sa, sb))
        out.append(y if x == 1 else x)
    return tuple(out)


def _bstrides(t, shape):
    """Strides of ``t`` viewed over the broadcast ``shape`` (0 where t is expanded)."""
    lead = len(shape) - len(t._shape)
    st = [0] * lead
    for s, have, stride in zip(shape[lead:], t._shape, t._strides):
        st.append(stride if have == s and s != 1 else 0)
    return st


def _ew1(op, a, alpha=0.0, out=None):
    if out is None:
        out = CudaTensor._new(a._shape, a._dtype)
    if a._numel == 0:
        return out
    if a._contig and out._contig:
        rt.api.ew_flat(op, a._code, a.ptr, None, None, out.ptr, a._numel, alpha)
    else:
        rt.api.ew(op, a._code, len(a._shape), i64arr(a._shape), a.ptr, i64arr(a._strides), None, None, None, None,
                  out.ptr, i64arr(out._strides), alpha)
    return out


def _ewn(op, operands, alpha=0.0, out=None):
    """out = op(*operands) with numpy broadcasting; operands are CudaTensors of one float dtype."""
    shape = operands[0]._shape
    same = True
    for t in operands[1:]:
        if t._shape != shape:
            same = False
            shape = _bshape(shape, t._shape)
    if out is None:
        out = CudaTensor._new(shape, operands[0]._dtype)
    elif out._shape != shape:
        raise ValueError("non-broadcastable output operand with shape %s doesn't match the broadcast shape %s"
                         % (out._shape, shape))
    n = out._numel
    if n == 0:
        return out
    code = operands[0]._code
    if same and out._contig and all(t._contig for t in operands):
        p = [t.ptr for t in operands] + [None, None]
        rt.api.ew_flat(op, code, p[0], p[1], p[2], out.ptr, n, alpha)
        return out
    args = []
    for t in operands:
        args += [t.ptr, i64arr(_bstrides(t, shape))]
    while len(args) < 6:
        args += [None, None]
    rt.api.ew(op, code, len(shape), i64arr(shape), *args, out.ptr, i64arr(out._strides), alpha)
    return out


def _assign(dst, val):
    """dst[...] = val with numpy broadcasting of ``val``; ``dst`` may be any strided view."""
    if val._code != dst._code:
        val = val.astype(dst._dtype)
    _written(dst)
    shp = dst._shape
    if _bshape(shp, val._shape) != shp:
        raise ValueError("could not broadcast input from shape %s into shape %s" % (val._shape, shp))
    if dst._numel:
        rt.api.cast(dst._code, dst._code, len(shp), i64arr(shp), val.ptr, i64arr(_bstrides(val, shp)),
                    dst.ptr, i64arr(dst._strides))


def _unbroadcast(g, shape):
    """Sum ``g`` down to ``shape`` (what Function._deliver would do later, done eagerly where it is cheaper)."""
    if g._shape == shape:
        return g
    lead = len(g._shape) - len(shape)
    axes = tuple(range(lead)) + tuple(lead + i for i, (x, y) in enumerate(zip(shape, g._shape[lead:])) if x != y)
    return _reduce(RED['SUM'], g, axes, True).reshape(*shape)


# ---------------------------------------------------------------------------------------------------
# transformations (cpu/ops.py:25-47, opencl/ops.py:9-36)
@CudaTensor.register_op()
@CudaTensor.register_op("T")
class transpose(Function):
    def forward(ctx, a, *axes):
        nd = len(a._shape)
        if len(axes) == 1 and isinstance(axes[0], (tuple, list)):
            axes = tuple(axes[0])
        if len(axes) == 0:
            axes = tuple(reversed(range(nd)))
        axes = tuple(ax % nd for ax in axes)
        assert sorted(axes) == list(range(nd)), "axes don't match tensor"
        ctx.save_for_backward(axes)
        out = a._view(tuple(a._shape[i] for i in axes), tuple(a._strides[i] for i in axes))
        if a._temp:
            # sole ownership of the buffer moves to the view
            a._temp, out._temp = False, True
        return out

    def backward(ctx, out_grad):
        axes, = ctx.get_saved_tensors()
        inv = [0] * len(axes)
        for i, j in enumerate(axes):
            inv[j] = i
        return out_grad.transpose(*inv)


def _reshape_strides(shape, strides, new_shape):
    """Strides that express ``new_shape`` over the same memory, or None if a copy is needed."""
    old = [(s, st) for s, st in zip(shape, strides) if s != 1]
    new_strides, oi = [], 0
    ni, n_new = 0, len(new_shape)
    if _prod(shape) == 0:
        return contiguous_strides(new_shape)
    while ni < n_new:
        if new_shape[ni] == 1:
            new_strides.append(1)
            ni += 1
            continue
        if oi >= len(old):
            return None
        # grow a group of old dims and a group of new dims until their sizes agree
        osz, ost_last = old[oi][0], old[oi][1]
        oj = oi + 1
        nsz, nj = new_shape[ni], ni + 1
        while osz != nsz:
            if osz < nsz:
                if oj >= len(old):
                    return None
                # old dims in one group must be mutually contiguous
                if old[oj - 1][1] != old[oj][0] * old[oj][1]:
                    return None
                osz *= old[oj][0]
                ost_last = old[oj][1]
                oj += 1
            else:
                if nj >= n_new:
                    return None
                nsz *= new_shape[nj]
                nj += 1
        # check contiguity inside the old group when it spans several dims
        for k in range(oi, oj - 1):
            if old[k][1] != old[k + 1][0] * old[k + 1][1]:
                return None
        # lay the new group's strides down from the innermost old stride
        acc = old[oj - 1][1]
        grp = []
        for k in range(nj - 1, ni - 1, -1):
            grp.append(acc)
            acc *= new_shape[k]
        new_strides.extend(reversed(grp))
        oi, ni = oj, nj
    return tuple(new_strides)


@CudaTensor.register_op()
class reshape(Function):
    def forward(ctx, a, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = tuple(int(s) for s in shape)
        if -1 in shape:
            known = -_prod(shape)
            shape = tuple(s if s != -1 else (a._numel // known if known else 0) for s in shape)
        assert _prod(shape) == a._numel, "cannot reshape tensor of size %d into shape %s" % (a._numel, shape)
        ctx.save_for_backward(a._shape)
        st = contiguous_strides(shape) if a._contig else _reshape_strides(a._shape, a._strides, shape)
        if st is None:
            a = a.copy()
            a._temp = False
            st = contiguous_strides(shape)
        out = a._view(shape, st)
        if a._temp:
            a._temp, out._temp = False, True
        return out

    def backward(ctx, out_grad):
        shape, = ctx.get_saved_tensors()
        return out_grad.reshape(*shape)


# ---------------------------------------------------------------------------------------------------
# elementwise arithmetic (cpu/ops.py:52-105, opencl/ops.py:40-114)
@CudaTensor.register_op()
class neg(Function):
    def forward(ctx, a):
        return _ew1(EW['NEG'], _float_like(a))

    def backward(ctx, out_grad):
        return _ew1(EW['NEG'], out_grad)


def _binary_forward(ctx, a, b, op, s_op, rs_op):
    """Common forward for add/sub/mul/div/pow: tensor-tensor, tensor-scalar and scalar-tensor."""
    if not isinstance(a, CudaTensor):
        if _is_scalar(a):
            b = _float_like(b)
            ctx.kind = ('rs', float(a))
            return _ew1(rs_op, b, float(a)) if rs_op is not None else None
        a = _as_tensor(a, b)
    if not isinstance(b, CudaTensor):
        if _is_scalar(b):
            a = _float_like(a)
            ctx.kind = ('s', float(b))
            return _ew1(s_op, a, float(b)) if s_op is not None else None
        b = _as_tensor(b, a)
    a, b = _promote(a, b)
    ctx.kind = ('tt', None)
    ctx.operands = (a, b)
    return _ewn(op, (a, b))


@CudaTensor.register_op()
class add(Function):
    def forward(ctx, a, b):
        return _binary_forward(ctx, a, b, EW['ADD'], EW['ADD_S'], EW['ADD_S'])

    def backward(ctx, out_grad):
        return out_grad, out_grad


@CudaTensor.register_op(overwrite=True)
@CudaTensor.register_op("__sub__", overwrite=True)
class sub(Function):
    def forward(ctx, a, b):
        if _is_scalar(b) and isinstance(a, CudaTensor):
            ctx.kind = ('s', float(b))
            return _ew1(EW['ADD_S'], _float_like(a), -float(b))
        return _binary_forward(ctx, a, b, EW['SUB'], None, EW['RSUB_S'])

    def backward(ctx, out_grad):
        return out_grad, _ew1(EW['NEG'], out_grad)


@CudaTensor.register_op("__rsub__", overwrite=True)
class rsub(Function):
    """a - b with b the tensor (ops.py:38-42 of the reference)."""
    def forward(ctx, b, a):
        if _is_scalar(a):
            return _ew1(EW['RSUB_S'], _float_like(b), float(a))
        a = _as_tensor(a, b)
        a, b = _promote(a, b)
        return _ewn(EW['SUB'], (a, b))

    def backward(ctx, out_grad):
        return _ew1(EW['NEG'], out_grad), out_grad


@CudaTensor.register_op()
class mul(Function):
    def forward(ctx, a, b):
        out = _binary_forward(ctx, a, b, EW['MUL'], EW['MUL_S'], EW['MUL_S'])
        if ctx.kind[0] == 'tt':
            for t in ctx.operands:
                t._mark_shared()
        return out

    def backward(ctx, out_grad):
        kind, s = ctx.kind
        if kind != 'tt':
            g = _ew1(EW['MUL_S'], out_grad, s)
            return g, g
        a, b = ctx.operands
        g = out_grad if out_grad._code == a._code else out_grad.astype(a._dtype)
        if a._shape == b._shape == g._shape and a._contig and b._contig and g._contig:
            da, db = CudaTensor._new(a._shape, a._dtype), CudaTensor._new(a._shape, a._dtype)
            if a._numel:
                rt.api.ew_bwd2_flat(0, a._code, a.ptr, b.ptr, g.ptr, da.ptr, db.ptr, a._numel)
            return da, db
        # broadcast operands: reduce eagerly so the big temporaries die here
        return _unbroadcast(_ewn(EW['MUL'], (g, b)), a._shape), _unbroadcast(_ewn(EW['MUL'], (a, g)), b._shape)


@CudaTensor.register_op(overwrite=True)
@CudaTensor.register_op("__truediv__", overwrite=True)
class div(Function):
    def forward(ctx, a, b):
        out = _binary_forward(ctx, a, b, EW['DIV'], EW['DIV_S'], EW['RDIV_S'])
        if ctx.kind[0] == 'tt':
            for t in ctx.operands:
                t._mark_shared()
        elif ctx.kind[0] == 'rs':
            ctx.operands = (_float_like(b),)
        return out

    def backward(ctx, out_grad):
        kind, s = ctx.kind
        if kind == 's':
            g = _ew1(EW['DIV_S'], out_grad, s)
            return g, g
        if kind == 'rs':
            g = _ewn(EW['RDIV_S_BWD'], (ctx.operands[0], out_grad), s)
            return g, g
        a, b = ctx.operands
        g = out_grad if out_grad._code == a._code else out_grad.astype(a._dtype)
        if a._shape == b._shape == g._shape and a._contig and b._contig and g._contig:
            da, db = CudaTensor._new(a._shape, a._dtype), CudaTensor._new(a._shape, a._dtype)
            if a._numel:
                rt.api.ew_bwd2_flat(1, a._code, a.ptr, b.ptr, g.ptr, da.ptr, db.ptr, a._numel)
            return da, db
        return (_unbroadcast(_ewn(EW['DIV'], (g, b)), a._shape),
                _unbroadcast(_ewn(EW['DIV_BWD_B'], (a, b, g)), b._shape))


@CudaTensor.register_op("__rtruediv__", overwrite=True)
class rdiv(Function):
    """a / b with b the tensor (ops.py:43-47 of the reference)."""
    def forward(ctx, b, a):
        b = _float_like(b)
        if _is_scalar(a):
            ctx.s, ctx.b = float(a), b
            b._mark_shared()
            return _ew1(EW['RDIV_S'], b, float(a))
        a = _as_tensor(a, b)
        a, b = _promote(a, b)
        ctx.s, ctx.a, ctx.b = None, a, b
        return _ewn(EW['DIV'], (a, b))

    def backward(ctx, out_grad):
        if ctx.s is not None:
            g = _ewn(EW['RDIV_S_BWD'], (ctx.b, out_grad), ctx.s)
            return g, g
        return (_unbroadcast(_ewn(EW['DIV_BWD_B'], (ctx.a, ctx.b, out_grad)), ctx.b._shape),
                _unbroadcast(_ewn(EW['DIV'], (out_grad, ctx.b)), ctx.a._shape))


@CudaTensor.register_op()
class pow(Function):
    def forward(ctx, a, b):
        if isinstance(a, CudaTensor) and _is_scalar(b):
            a = _float_like(a)
            e = float(b)
            ctx.kind = ('s', e)
            ctx.operands = (a,)
            a._mark_shared()
            # numpy evaluates these exponents with exact kernels (square / sqrt / reciprocal)
            if e == 2.0:
                return _ewn(EW['MUL'], (a, a))
            if e == 0.5:
                return _ew1(EW['SQRT'], a)
            if e == -1.0:
                return _ew1(EW['RDIV_S'], a, 1.0)
            if e == 1.0:
                return _ew1(EW['COPY'], a)
            return _ew1(EW['POW_S'], a, e)
        if _is_scalar(a):
            b = _float_like(b)
            y = _ew1(EW['RPOW_S'], b, float(a))
            ctx.kind = ('rs', float(a))
            ctx.operands = (y,)
            y._mark_shared()
            return y
        out = _binary_forward(ctx, a, b, EW['POW'], None, None)
        out._mark_shared()
        ctx.operands = ctx.operands + (out,)
        return out

    def backward(ctx, out_grad):
        kind, s = ctx.kind
        if kind == 's':
            # the reference also evaluates ln(a) for the (non-tensor) exponent and discards it
            g = _ewn(EW['POW_S_BWD'], (ctx.operands[0], out_grad), s)
            return g, g
        if kind == 'rs':
            g = _ewn(EW['RPOW_S_BWD'], (ctx.operands[0], out_grad), s)
            return g, g
        a, b, y = ctx.operands
        g = out_grad if out_grad._code == a._code else out_grad.astype(a._dtype)
        return (_unbroadcast(_ewn(EW['POW_BWD_A'], (a, b, g)), a._shape),
                _unbroadcast(_ewn(EW['POW_BWD_B'], (a, y, g)), b._shape))


# ---------------------------------------------------------------------------------------------------
# in-place operators (cpu/ops.py:120-157, opencl/ops.py:136-177): mutate the operand's storage and
# hand back a new wrapper over the same memory, exactly as the reference adapters do.
def _inplace(op, s_op, negate=False):
    class _Op(Function):
        def forward(ctx, t, other):
            rt.side_join_if_written(t)
            _written(t)
            if _is_scalar(other):
                v = -float(other) if negate else float(other)
                if t._code not in (rt.F32, rt.F64):
                    raise TypeError("in-place arithmetic needs a float tensor")
                _ew1(s_op, t, v, out=t)
            else:
                other = _as_tensor(other, t)
                if other._code != t._code:
                    other = other.astype(t._dtype)
                _ewn(op, (t, other), out=t)
            return t._view(t._shape, t._strides)
    return _Op


for _name, _op, _sop, _neg in (('__iadd__', EW['ADD'], EW['ADD_S'], False),
                               ('__isub__', EW['SUB'], EW['ADD_S'], True),
                               ('__imul__', EW['MUL'], EW['MUL_S'], False),
                               ('__itruediv__', EW['DIV'], EW['DIV_S'], False)):
    _cls = _inplace(_op, _sop, _neg)
    _cls.__name__ = _name.strip('_')
    CudaTensor.register_op(_name, _cls, overwrite=True)


@CudaTensor.register_op()
class fill(Function):
    def forward(ctx, t, val):
        _written(t)
        t._fill_value(val)
        return t._view(t._shape, t._strides)


# ---------------------------------------------------------------------------------------------------
# unary math (cpu/ops.py:158-229, opencl/ops.py:182-288)
def _unary(name, fwd_op, bwd_op, save_output):
    class _Op(Function):
        def forward(ctx, t):
            t = _float_like(t)
            y = _ew1(fwd_op, t)
            ctx.save_for_backward(y if save_output else t)
            return y

        def backward(ctx, out_grad):
            s, = ctx.get_saved_tensors()
            g = out_grad if out_grad._code == s._code else out_grad.astype(s._dtype)
            return _ewn(bwd_op, (s, g))
    _Op.__name__ = name
    return _Op


for _name, _f, _b, _so, _ow in (('sin', 'SIN', 'SIN_BWD', False, False), ('cos', 'COS', 'COS_BWD', False, False),
                                ('exp', 'EXP', 'MUL', True, False), ('log', 'LOG', 'LOG_BWD', False, False),
                                ('sigmoid', 'SIGMOID', 'SIGMOID_BWD', True, True),
                                ('tanh', 'TANH', 'TANH_BWD', True, True),
                                ('relu', 'RELU', 'RELU_BWD', False, False),
                                ('gelu', 'GELU', 'GELU_BWD', False, False)):
    CudaTensor.register_op(_name, _unary(_name, EW[_f], EW[_b], _so), overwrite=_ow)


# ---------------------------------------------------------------------------------------------------
# reductions (cpu/ops.py:260-293, opencl/ops.py:335-400)
def _norm_axes(axis, nd):
    if axis is None:
        return tuple(range(nd))
    if isinstance(axis, (int, np.integer)):
        axis = (int(axis),)
    return tuple(sorted(set(int(a) % nd for a in axis))) if nd else ()


def _reduce(op, x, axes, keepdims, scale=1.0):
    """Reduce contiguous runs of axes with one (outer, reduce, inner) kernel each."""
    x = _float_like(x)
    nd = len(x._shape)
    axes = _norm_axes(axes, nd)
    full_shape = x._shape
    if x._numel == 0:
        raise ValueError("zero-size array to reduction operation")
    if not x._contig and nd >= 2 and axes == tuple(range(nd - 1)) and x._strides[-1] == 1:
        # rows with a padded pitch (a GEMM result written for TMA alignment) reduced over all leading dims:
        # the column-reduce kernel reads them in place
        rows = _prod(full_shape[:-1])
        st = _reshape_strides(full_shape, x._strides, (rows, full_shape[-1]))
        if st is not None and st[1] == 1 and st[0] >= full_shape[-1]:
            kshape = (1,) * (nd - 1) + (full_shape[-1],)
            out = CudaTensor._new(kshape if keepdims else (full_shape[-1],), x._dtype)
            rt.api.reduce_pitched(op, x._code, x.ptr, out.ptr, 1, rows, full_shape[-1], st[0], scale, 0)
            return out
    cur = x.contiguous()
    shape = list(full_shape)
    # group adjacent axes, reduce from the last group to the first so earlier dims keep their place
    groups, run = [], []
    for a in axes:
        if run and a == run[-1] + 1:
            run.append(a)
        else:
            if run:
                groups.append(run)
            run = [a]
    if run:
        groups.append(run)
    for gi, grp in enumerate(reversed(groups)):
        lo, hi = grp[0], grp[-1] + 1
        outer, red, inner = _prod(shape[:lo]), _prod(shape[lo:hi]), _prod(shape[hi:])
        for a in grp:
            shape[a] = 1
        out = CudaTensor._new(tuple(shape), cur._dtype)
        last = gi == len(groups) - 1
        rt.api.reduce(op, cur._code, cur.ptr, out.ptr, outer, red, inner, scale if last else 1.0)
        cur = out
    if not groups:
        cur = x.copy() if scale == 1.0 else _ew1(EW['MUL_S'], x, scale)
    if not keepdims:
        final = tuple(s for i, s in enumerate(full_shape) if i not in axes)
        temp = cur._temp
        cur = cur._view(final, contiguous_strides(final))
        cur._temp = temp
    return cur


def _keep_shape(shape, axes):
    return tuple(1 if i in axes else s for i, s in enumerate(shape))


def _extreme(name, red_op):
    class _Op(Function):
        def forward(ctx, x, axis=None, keepdims=False):
            x = _float_like(x)
            axes = _norm_axes(axis, len(x._shape))
            val = _reduce(red_op, x, axes, True)
            val._temp = False
            ctx.save_for_backward(x, val, axes)
            if keepdims:
                return val._view(val._shape, val._strides)
            final = tuple(s for i, s in enumerate(x._shape) if i not in axes)
            return val._view(final, contiguous_strides(final))

        def backward(ctx, out_grad):
            x, val, axes = ctx.get_saved_tensors()
            g = out_grad._view(val._shape, contiguous_strides(val._shape)) if out_grad._contig \
                else out_grad.contiguous()._view(val._shape, contiguous_strides(val._shape))
            if g._code != x._code:
                g = g.astype(x._dtype)
            return _ewn(EW['EQ_MASK_MUL'], (x, val, g))
    _Op.__name__ = name
    return _Op


CudaTensor.register_op('max', _extreme('max', RED['MAX']))
CudaTensor.register_op('min', _extreme('min', RED['MIN']))


@CudaTensor.register_op()
class sum(Function):
    def forward(ctx, t, axis=None, keepdims=False):
        axes = _norm_axes(axis, len(t._shape))
        ctx.save_for_backward(t._shape, axes)
        return _reduce(RED['SUM'], t, axes, keepdims)

    def backward(ctx, out_grad):
        # broadcast back as a zero-stride view (as opencl/ops.py:344-368); materialised on accumulation
        shape, axes = ctx.get_saved_tensors()
        kshape = _keep_shape(shape, axes)
        g = out_grad if out_grad._contig else out_grad.contiguous()
        kst = contiguous_strides(kshape)
        st = tuple(0 if (i in axes or kshape[i] == 1) else kst[i] for i in range(len(shape)))
        return g._view(shape, st)


@CudaTensor.register_op(overwrite=True)
class mean(Function):
    """sum * (numel_out / numel_in) in one reduction (generic form: ops.py:71-75 of the reference)."""
    def forward(ctx, t, axis=None, keepdims=False):
        axes = _norm_axes(axis, len(t._shape))
        n_red = _prod(tuple(t._shape[a] for a in axes))
        ctx.save_for_backward(t._shape, axes, 1.0 / n_red)
        return _reduce(RED['SUM'], t, axes, keepdims, scale=1.0 / n_red)

    def backward(ctx, out_grad):
        shape, axes, scale = ctx.get_saved_tensors()
        kshape = _keep_shape(shape, axes)
        g = _ew1(EW['MUL_S'], out_grad if out_grad._contig else out_grad.contiguous(), scale)
        kst = contiguous_strides(kshape)
        st = tuple(0 if (i in axes or kshape[i] == 1) else kst[i] for i in range(len(shape)))
        return g._view(shape, st)


# ---------------------------------------------------------------------------------------------------
# matmul (cpu/ops.py:107-116, opencl/ops.py:116-132 + kernels.py:201-337)
def _collapse_batch(shape, strides_list):
    """Merge batch dims that are jointly mergeable for every operand.  Returns (shape, [strides...])."""
    dims = [(s, [st[i] for st in strides_list]) for i, s in enumerate(shape) if s != 1]
    if not dims:
        return [], [[] for _ in strides_list]
    out = [dims[0]]
    for s, sts in dims[1:]:
        ps, psts = out[-1]
        if all(p == s * q for p, q in zip(psts, sts)):
            out[-1] = (ps * s, sts)
        else:
            out.append((s, sts))
    return [d[0] for d in out], [[d[1][k] for d in out] for k in range(len(strides_list))]


def _gemm(a, b, out=None, bias=None, accumulate=False, mode=None, cls='F'):
    """out[..., M, N] = a[..., M, K] @ b[..., K, N] for float CudaTensors of >= 2 dims (views welcome)."""
    M, K = a._shape[-2:]
    K2, N = b._shape[-2:]
    if K != K2:
        raise ValueError("matmul: shapes %s and %s are not aligned" % (a._shape, b._shape))
    ba, bb = a._shape[:-2], b._shape[:-2]
    bshape = _bshape(ba, bb)
    mode = _mode_for(cls, mode)
    if out is None:
        use_mode = mode
        if use_mode != rt.GEMM_FP32_SIMT and N % 4 != 0 and N >= 64 and a._code == rt.F32:
            # TMA needs 16-byte aligned row strides: pad the leading dimension and hand back a view
            ld = (N + 31) // 32 * 32
            full = CudaTensor._new(bshape + (M, ld), a._dtype)
            out = full._view(bshape + (M, N), full._strides)
        else:
            out = CudaTensor._new(bshape + (M, N), a._dtype)
    sa = _bstrides(a._view(ba, a._strides[:-2]), bshape) if bshape else []
    sb = _bstrides(b._view(bb, b._strides[:-2]), bshape) if bshape else []
    sc = list(out._strides[:-2])
    cshape, (csa, csb, csc) = _collapse_batch(bshape, [sa, sb, sc])
    if len(cshape) > 2:
        # rare: more than two irreducible batch dims -> densify the operands once
        a2 = a.contiguous() if not a._contig else a
        b2 = b.contiguous() if not b._contig else b
        if len(_collapse_batch(bshape, [_bstrides(a2._view(ba, a2._strides[:-2]), bshape),
                                       _bstrides(b2._view(bb, b2._strides[:-2]), bshape), sc])[0]) > 2:
            a2 = _ew1(EW['COPY'], a._view(bshape + (M, K), tuple(sa) + a._strides[-2:]))
            b2 = _ew1(EW['COPY'], b._view(bshape + (K, N), tuple(sb) + b._strides[-2:]))
        return _gemm(a2, b2, out, bias, accumulate, mode, cls)
    while len(cshape) < 2:
        cshape.insert(0, 1)
        csa.insert(0, 0)
        csb.insert(0, 0)
        csc.insert(0, 0)
    d = rt.GemmDesc(M, N, K, cshape[0], cshape[1],
                    csa[0], csa[1], a._strides[-2], a._strides[-1],
                    csb[0], csb[1], b._strides[-2], b._strides[-1],
                    csc[0], csc[1], out._strides[-2], out._strides[-1])
    if out._numel:
        _written(out)
        if K == 0:
            out._fill_value(0)
        else:
            _launch_gemm(mode, a._code, d, a, b, out, bias, accumulate)
    return out


def _with_shape(t, shape):
    """Re-view a fresh GEMM result under ``shape`` (same memory; keeps the sole-owner flag)."""
    st = contiguous_strides(shape) if t._contig else _reshape_strides(t._shape, t._strides, shape)
    v = t._view(shape, st)
    v._temp = t._temp
    return v


def _fold_rows(t):
    """View (..., M, K) as (prod(...)*M, K) without copying when the leading dims are mergeable, else copy."""
    if len(t._shape) == 2:
        return t
    rows = _prod(t._shape[:-1])
    st = _reshape_strides(t._shape, t._strides, (rows, t._shape[-1]))
    if st is None:
        t = t.copy()
        st = contiguous_strides((rows, t._shape[-1]))
    return t._view((rows, t._shape[-1]), st)


def _empty_like_layout(ref):
    """Uninitialised tensor of ref's shape whose memory order follows ref's strides (None if ref broadcasts)."""
    shape, strides = ref._shape, ref._strides
    if ref._contig:
        return CudaTensor._new(shape, ref._dtype)
    if any(st <= 0 and s > 1 for s, st in zip(shape, strides)):
        return None
    order = sorted(range(len(shape)), key=lambda i: (-strides[i] if shape[i] > 1 else 0, i))
    dense = contiguous_strides(tuple(shape[i] for i in order))
    st = [0] * len(shape)
    for pos, i in enumerate(order):
        st[i] = dense[pos]
    base = CudaTensor._new(tuple(shape[i] for i in order), ref._dtype)
    out = base._view(shape, tuple(st))
    out._temp = True
    return out


def _gemm_like(ref, x, y, cls='F'):
    """x @ y written in the memory layout of ``ref`` (the operand this is the gradient of), so that the
    transpose / reshape backward that follows stays a view instead of a strided copy."""
    if ref._contig or ref._shape[:-2] != _bshape(x._shape[:-2], y._shape[:-2]) or ref._code != x._code:
        return _gemm(x, y, cls=cls)
    out = _empty_like_layout(ref)
    if out is None:
        return _gemm(x, y, cls=cls)
    if out._strides[-1] == 1:
        return _gemm(x, y, out=out, cls=cls)
    if out._strides[-2] == 1:
        # column-major result: compute the transposed product into the transposed view
        _gemm(_swap_last(y), _swap_last(x), out=_swap_last(out), cls=cls)
        return out
    return _gemm(x, y, cls=cls)


def _swap_last(t):
    return t._view(t._shape[:-2] + (t._shape[-1], t._shape[-2]), t._strides[:-2] + (t._strides[-1], t._strides[-2]))


@CudaTensor.register_op()
@CudaTensor.register_op("__matmul__")
class dot(Function):
    def forward(ctx, a, b, out_layout=None):
        """``out_layout='rhs'``: lay the result out in memory like ``b`` (same logical shape required), e.g.
        attention's P @ V lands directly in the (batch, seq, head, dim) order its head-merge expects."""
        if not isinstance(b, CudaTensor):
            b = _as_tensor(b, a)
        a, b = _promote(a, b)
        ctx.va, ctx.vb = len(a._shape) == 1, len(b._shape) == 1
        if ctx.va:
            a = a._view((1,) + a._shape, (0,) + a._strides)
        if ctx.vb:
            b = b._view(b._shape + (1,), b._strides + (0,))
        a._mark_shared()
        b._mark_shared()
        ctx.save_for_backward(a, b)
        if len(b._shape) == 2 and len(a._shape) > 2:
            # (batch.., M, K) @ (K, N): one GEMM with M' = batch*M
            a2 = _fold_rows(a)
            out = _with_shape(_gemm(a2, b), a._shape[:-1] + (b._shape[-1],))
        elif out_layout == 'rhs' and not b._contig and b._shape == a._shape[:-1] + b._shape[-1:] \
                and a._shape[:-2] == b._shape[:-2]:
            out = _empty_like_layout(b)
            out = _gemm(a, b, out=out) if out is not None and out._strides[-1] == 1 else _gemm(a, b)
        else:
            out = _gemm(a, b)
        if ctx.va or ctx.vb:
            shp = list(out._shape)
            if ctx.vb:
                shp.pop(-1)
            if ctx.va:
                shp.pop(-2 if not ctx.vb else -1)
            out = _with_shape(out, tuple(shp))
        return out

    def backward(ctx, out_grad):
        a, b = ctx.get_saved_tensors()
        g = out_grad if out_grad._code == a._code else out_grad.astype(a._dtype)
        if ctx.va or ctx.vb:
            shp = list(g._shape)
            if ctx.va:
                shp.insert(len(shp) - (0 if ctx.vb else 1), 1)
            if ctx.vb:
                shp.append(1)
            g = g.contiguous()._view(tuple(shp), None)
        if len(b._shape) == 2 and len(a._shape) > 2:
            # dA = g @ b^T per row; dB = A2d^T @ g2d : the batch sum is folded into the GEMM's K dim
            g2, a2 = _fold_rows(g), _fold_rows(a)
            da = _with_shape(_gemm(g2, _swap_last(b), cls='X'), a._shape)
            db = _gemm(_swap_last(a2), g2, cls='W')
        else:
            da = _gemm_like(a, g, _swap_last(b), cls='X')
            db = _gemm_like(b, _swap_last(a), g, cls='X')
            da, db = _unbroadcast(da, a._shape), _unbroadcast(db, b._shape)
        if ctx.va:
            da = da.reshape(da._shape[-1])
        if ctx.vb:
            db = db.reshape(*db._shape[:-1])
        return da, db


@CudaTensor.register_op()
class linear(Function):
    """y = x @ W^T (+ b) with W stored (out, in) as nn.Linear keeps it (nn.py:90-96 of the reference).

    One GEMM forward (bias fused into the epilogue) and two backward: dX = dY @ W and
    dW = dY^T @ X written directly in W's layout, with the batch/sequence dims folded into the
    reduction (the reference computes a batched dW and then sums it over the batch).
    """
    def forward(ctx, x, weight, bias=None):
        x = _float_like(x)
        x2 = _fold_rows(x)
        x2._mark_shared()
        ctx.save_for_backward(x2, weight, x._shape, bias is not None)
        return _with_shape(_gemm(x2, _swap_last(weight), bias=bias), x._shape[:-1] + (weight._shape[0],))

    def backward(ctx, out_grad):
        x2, weight, xshape, has_bias = ctx.get_saved_tensors()
        g2 = _fold_rows(out_grad)
        xin = ctx._parents[0]
        xg = _direct_grad(xin, g2._code) if isinstance(xin, CudaTensor) and xin._shape == tuple(xshape) else None
        if xg is not None:
            # x already holds a gradient from another consumer (residual branch): dX is reduce-added into it by
            # the GEMM epilogue instead of being materialised and added by a separate pass
            rt.side_join_if_written(xg)
            _gemm(g2, weight, out=_fold_rows(xg), accumulate=True, cls='X')
            dx = Function.ACCUMULATED
        else:
            dx = _with_shape(_gemm(g2, weight, cls='X'), xshape)
        wg = _direct_grad(weight, g2._code)
        bias = ctx._parents[2] if has_bias else None
        bg = _direct_grad(bias, g2._code) if has_bias else None
        if bg is not None and not (bg._shape == (g2._shape[1],) and g2._strides[1] == 1 and g2._shape[1] > 1):
            bg = None
        if wg is not None and (bg is not None or not has_bias):
            # dW += dY^T X and db += column sums of dY go straight into the existing gradients (the optimizer's
            # arena: no temporary, no add pass).  Nothing else in backward waits for them, so they are issued on
            # the side stream and overlap the dX chain.
            with rt.side_stream(g2, x2, writes=(wg,) if bg is None else (wg, bg)):
                _gemm(_swap_last(g2), x2, out=wg, accumulate=True, cls='W')
                if bg is not None:
                    _bias_grad(g2._code, g2.ptr, bg.ptr, g2._shape[0], g2._shape[1], g2._strides[0], src=g2)
            return (dx, Function.ACCUMULATED, Function.ACCUMULATED) if has_bias else (dx, Function.ACCUMULATED)
        if wg is not None:
            rt.side_join_if_written(wg)       # a weight shared with another layer may have side-stream writes pending
            _gemm(_swap_last(g2), x2, out=wg, accumulate=True, cls='W')
            dw = Function.ACCUMULATED
        else:
            dw = _gemm(_swap_last(g2), x2, cls='W')
        if has_bias:
            return dx, dw, _reduce(RED['SUM'], g2, (0,), False)
        return dx, dw


def _gemm_grouped(As, Bs, outs, biases=None, accumulate=False, cls='F'):
    """outs[g] (+)= As[g] @ Bs[g] (+ biases[g]) for up to 4 problems of identical shape and strides, as one
    launch.  Naming the same tensor in every ``outs`` slot makes it one K-concatenated product
    out = sum_g As[g] @ Bs[g].  Operands: 2-D, or N-D with batch dims that collapse to <= 2 strides."""
    a, b, out = As[0], Bs[0], outs[0]
    for x, y, o in zip(As, Bs, outs):
        assert x._shape == a._shape and x._strides == a._strides and x._code == a._code
        assert y._shape == b._shape and y._strides == b._strides and y._code == a._code
        assert o._shape == out._shape and o._strides == out._strides and o._code == a._code
    M, K = a._shape[-2:]
    N = b._shape[-1]
    assert b._shape[-2] == K and out._shape[-2:] == (M, N) and a._shape[:-2] == b._shape[:-2] == out._shape[:-2]
    bshape = a._shape[:-2]
    cshape, (csa, csb, csc) = _collapse_batch(bshape, [list(a._strides[:-2]), list(b._strides[:-2]),
                                                       list(out._strides[:-2])])
    assert len(cshape) <= 2, "grouped matmul: batch dims must collapse to two"
    while len(cshape) < 2:
        cshape.insert(0, 1)
        csa.insert(0, 0)
        csb.insert(0, 0)
        csc.insert(0, 0)
    d = rt.GemmDesc(M, N, K, cshape[0], cshape[1],
                    csa[0], csa[1], a._strides[-2], a._strides[-1],
                    csb[0], csb[1], b._strides[-2], b._strides[-1],
                    csc[0], csc[1], out._strides[-2], out._strides[-1])
    n = len(As)
    arr = C.c_void_p * n
    for o in outs:
        _written(o)
    mode, code = _mode_for(cls), a._code
    pa, pb = [t.ptr for t in As], [t.ptr for t in Bs]
    if mode == rt.GEMM_BF16_TC:
        ok = code == rt.F32 and not (accumulate and biases is not None) and \
            rt.api.gemm_tc_supported(mode, rt.BF16, C.byref(d))
        if ok:
            qa, qb = [_bf16_ptr(t) for t in As], [_bf16_ptr(t) for t in Bs]
            ok = not any((x | y | o.ptr) & 15 for x, y, o in zip(qa, qb, outs))
        if ok:
            pa, pb, code = qa, qb, rt.BF16
        else:
            mode = rt.GEMM_FP32_SIMT
    rt.api.gemm_grouped(mode, code, C.byref(d), n, arr(*pa), arr(*pb), arr(*[t.ptr for t in outs]),
                        arr(*[t.ptr for t in biases]) if biases is not None else None, 1 if accumulate else 0)
    return outs


def _gemm_epilogue(a, b, out, bias, epi, aux, cls='F'):
    """One plain 2-D product with an activation fused into the GEMM epilogue (lg_gemm_epilogue):
    epi 1: out = a @ b + bias and aux = gelu(out);   epi 2: out = (a @ b) * gelu'(aux)."""
    M, K = a._shape
    N = b._shape[1]
    assert b._shape[0] == K and out._shape == (M, N) == aux._shape and out._strides[1] == 1 == aux._strides[1]
    d = rt.GemmDesc(M, N, K, 1, 1, 0, 0, a._strides[0], a._strides[1], 0, 0, b._strides[0], b._strides[1],
                    0, 0, out._strides[0], 1)
    _written(out)
    if epi == 1:
        _written(aux)
    mode, code, pa, pb = _mode_for(cls), a._code, a.ptr, b.ptr     # (epi 2: ``bias`` receives the result's column sums)
    if mode == rt.GEMM_BF16_TC:
        ok = code == rt.F32 and rt.api.gemm_tc_supported(mode, rt.BF16, C.byref(d))
        if ok:
            qa, qb = _bf16_ptr(a), _bf16_ptr(b)
            ok = not ((qa | qb | out.ptr) & 15)
        if ok:
            pa, pb, code = qa, qb, rt.BF16
        else:
            mode = rt.GEMM_FP32_SIMT
    rt.api.gemm_epilogue(mode, code, C.byref(d), pa, pb, out.ptr,
                         bias.ptr if bias is not None else None, epi, aux.ptr, aux._strides[0], 0.0)
    return out


def _attention_gemm(a, b, out, epi, alpha, aux=None):
    """Batched (batch, heads, M, K) @ (batch, heads, K, N) into the dense (batch, heads, M, N) ``out`` with a row
    epilogue (lg_gemm_epilogue 3: softmax(alpha * product); 4: alpha * aux * (product - rowsum(aux * product))).
    Returns False -- nothing launched -- when the tensor-core row epilogue cannot take the problem."""
    if _matmul_mode == rt.GEMM_FP32_SIMT or 'attn_epi' in _DISABLED:
        return False
    B0, B1, M, K = a._shape
    N = b._shape[-1]
    if N > 128 or N % 4 or not out._contig or (aux is not None and not aux._contig) or a._code != rt.F32:
        return False
    d = rt.GemmDesc(M, N, K, B0, B1, a._strides[0], a._strides[1], a._strides[2], a._strides[3],
                    b._strides[0], b._strides[1], b._strides[2], b._strides[3],
                    out._strides[0], out._strides[1], out._strides[2], 1)
    mode = _mode_for('A')
    code, pa, pb = a._code, a.ptr, b.ptr
    if mode == rt.GEMM_BF16_TC:
        code = rt.BF16
    if not rt.api.gemm_tc_supported(mode, code, C.byref(d)):
        return False
    if code == rt.BF16:
        pa, pb = _bf16_ptr(a), _bf16_ptr(b)
        if (pa | pb) & 15:
            return False
    _written(out)
    rt.api.gemm_epilogue(mode, code, C.byref(d), pa, pb, out.ptr, None, epi,
                         aux.ptr if aux is not None else None, N, float(alpha))
    return True


@CudaTensor.register_op()
class mlp_gelu(Function):
    """y = gelu(x W1^T + b1) W2^T + b2 -- the feed-forward block of BertLayer (examples/bert.py:150-153 of the
    reference: output.dense(gelu(intermediate.dense(x)))) as one graph node.

    Forward: the first GEMM's epilogue adds the bias and writes both the pre-activation h (kept for backward)
    and gelu(h); backward: the GEMM that forms d(gelu(h)) = dY W2 multiplies by gelu'(h) in its epilogue, so the
    (rows x intermediate) activation gradient is never written un-multiplied; weight / bias gradients go to the
    side stream and into the optimizer's arena; dX is added into x's gradient when the residual branch already
    delivered one.
    """
    def forward(ctx, x, w1, b1, w2, b2):
        x = _float_like(x)
        x2 = _fold_rows(x)
        if not x2._contig and x2._strides[1] != 1:
            x2 = x2.contiguous()
        x2._mark_shared()
        rows, F = x2._shape[0], w1._shape[0]
        h = CudaTensor._new((rows, F), x._dtype)
        act = CudaTensor._new((rows, F), x._dtype)
        _gemm_epilogue(x2, _swap_last(w1), h, b1, 1, act)
        y = _gemm(act, _swap_last(w2), bias=b2)
        h._temp = act._temp = False
        ctx.save_for_backward(x2, h, act, x._shape)
        return _with_shape(y, x._shape[:-1] + (w2._shape[0],))

    def backward(ctx, out_grad):
        x2, h, act, xshape = ctx.get_saved_tensors()
        xin, w1, b1, w2, b2 = ctx._parents[:5]
        g2 = _fold_rows(out_grad)
        if g2._code != x2._code:
            g2 = g2.astype(x2._dtype)
        if g2._strides[1] != 1:
            g2 = g2.contiguous()
        code = x2._code
        dh = CudaTensor._new(h._shape, h._dtype)
        grads = [_direct_grad(p, code) for p in (w1, b1, w2, b2)]
        fused_ok = all(g is not None for g in grads) and min(w1._shape[0], w2._shape[0]) > 1
        # dh = (dY W2) * gelu'(h); with the gradient arena in place the same epilogue also adds the column sums of dh
        # -- the gradient of b1 -- to it, instead of a separate pass over the (rows x intermediate) matrix
        b1_in_epilogue = fused_ok and 'bias_epi' not in _DISABLED
        if b1_in_epilogue:
            rt.side_join_if_written(grads[1])
        _gemm_epilogue(g2, w2, dh, grads[1] if b1_in_epilogue else None, 2, h, cls='X')
        xg = _direct_grad(xin, code) if isinstance(xin, CudaTensor) and xin._shape == tuple(xshape) else None
        if xg is not None:
            rt.side_join_if_written(xg)
            _gemm(dh, w1, out=_fold_rows(xg), accumulate=True, cls='X')
            dx = Function.ACCUMULATED
        else:
            dx = _with_shape(_gemm(dh, w1, cls='X'), xshape)
        if fused_ok:
            w1g, b1g, w2g, b2g = grads
            with rt.side_stream(g2, dh, act, x2, writes=tuple(grads)):
                _gemm(_swap_last(g2), act, out=w2g, accumulate=True, cls='W')
                _bias_grad(code, g2.ptr, b2g.ptr, g2._shape[0], g2._shape[1], g2._strides[0], src=g2)
                _gemm(_swap_last(dh), x2, out=w1g, accumulate=True, cls='W')
                if not b1_in_epilogue:
                    _bias_grad(code, dh.ptr, b1g.ptr, dh._shape[0], dh._shape[1], dh._strides[0])
            return (dx,) + (Function.ACCUMULATED,) * 4
        return (dx, _gemm(_swap_last(dh), x2, cls='W'), _reduce(RED['SUM'], dh, (0,), False),
                _gemm(_swap_last(g2), act, cls='W'), _reduce(RED['SUM'], g2, (0,), False))


def _bias_grad(code, g_ptr, out_ptr, rows, cols, ld, src=None):
    """out[cols] += column sums of the (rows, cols) gradient at g_ptr (row pitch ld): a bias gradient added straight
    into its slot of the gradient arena.  When the kernel that produced the gradient (``src``: LayerNorm backward)
    left its column sums next to it, they are added instead of reading the matrix back.
    (LG_DISABLE=bias_grad skips it all: a timing experiment, the result is wrong.)"""
    if 'bias_grad' in _DISABLED:
        return
    note = _root(src)._colsum if src is not None else None
    if note is not None and note[1:] == (g_ptr, rows, cols, ld) and note[0]._code == code:
        rt.api.ew_flat(EW['ADD'], code, out_ptr, note[0].ptr, None, out_ptr, cols, 0.0)
        return
    rt.api.reduce_pitched(RED['SUM'], code, g_ptr, out_ptr, 1, rows, cols, ld, 1.0, 1)


def _fused_attention_ok(code, s, dh):
    """The one-kernel attention runs tf32 products: used in the tensor-core modes (in bf16 mode when the attention
    class stays tf32), never in the exact-fp32 mode."""
    if 'attn_fused' in _DISABLED or _mode_for('A') != rt.GEMM_TF32_TC:
        return False
    return bool(rt.api.attention_supported(code, s, dh))


def _direct_grad(p, code):
    """The parameter's existing gradient buffer when a kernel may add into it in place, else None."""
    if 'dx_acc' in _DISABLED and p.ctx is not None:
        return None
    g = p.grad if isinstance(p, CudaTensor) and p.requires_grad else None
    if g is not None and g._contig and g._code == code == rt.F32 and g._shape == p._shape:
        return g
    return None


@CudaTensor.register_op()
class self_attention(Function):
    """Multi-head self-attention without mask, heads merged: (b, s, H) -> (b, s, H)

        Q, K, V = x Wq^T + bq, x Wk^T + bk, x Wv^T + bv ;  out = softmax(Q K^T / sqrt(d)) V   per head

    i.e. BertSelfAttention.forward of the reference (examples/bert.py:60-93) as one graph node.  The three
    projections are one grouped GEMM into a stacked (3, b*s, H) buffer whose per-head views feed the batched
    score / context GEMMs without a copy; backward writes dQ, dK, dV into one stacked buffer, forms dX as ONE
    K-concatenated GEMM (sum of the three dY_g W_g accumulated in tensor memory) and the three dW as one
    grouped GEMM accumulating into the optimizer's gradient arena.
    """
    def forward(ctx, x, wq, bq, wk, bk, wv, bv, heads=1):
        x = _float_like(x)
        assert len(x._shape) == 3, "self_attention expects (batch, seq, hidden)"
        b, s, H = x._shape
        assert H % heads == 0 and wq._shape == wk._shape == wv._shape == (H, H)
        dh = H // heads
        rows = b * s
        x2 = _fold_rows(x)
        x2._mark_shared()
        qkv = CudaTensor._new((3, rows, H), x._dtype)
        parts = [qkv._view((rows, H), (H, 1), g * rows * H) for g in range(3)]
        _gemm_grouped([x2] * 3, [_swap_last(w) for w in (wq, wk, wv)], parts, [bq, bk, bv])
        hv = (b, heads, s, dh), (s * H, dh, H, 1)                      # per-head view of a (rows, H) matrix
        q, k, v = (qkv._view(hv[0], hv[1], g * rows * H) for g in range(3))
        scale = 1.0 / float(np.sqrt(dh))
        if _fused_attention_ok(x._code, s, dh):
            # scores, softmax and context in one kernel per (batch, head): the probabilities never leave the chip
            out = CudaTensor._new((b, s, H), x._dtype)
            lse = CudaTensor._new((b * heads * s,), x._dtype)
            rt.api.attention_fwd(x._code, qkv.ptr, b, s, heads, dh, scale, out.ptr, lse.ptr)
            qkv._temp = out._temp = lse._temp = False
            ctx.save_for_backward(x2, qkv, (out, lse), (b, s, H, heads, scale), x._shape)
            return out._view(out._shape, out._strides)
        probs = CudaTensor._new((b, heads, s, s), x._dtype)
        if not _attention_gemm(q, _swap_last(k), probs, 3, scale):      # softmax fused into the score GEMM's epilogue
            scores = _gemm(q, _swap_last(k), cls='A')
            rt.api.softmax_fwd(x._code, scores.ptr, probs.ptr, scores._numel // s, s, scale)
            del scores
        out = CudaTensor._new((b, s, H), x._dtype)
        _gemm(probs, v, out=out._view(hv[0], hv[1]), cls='A')
        probs._temp = qkv._temp = False
        ctx.save_for_backward(x2, qkv, probs, (b, s, H, heads, scale), x._shape)
        return out

    def backward(ctx, out_grad):
        x2, qkv, probs, (b, s, H, heads, scale), xshape = ctx.get_saved_tensors()
        wq, bq, wk, bk, wv, bv = ctx._parents[1:7]
        dh, rows = H // heads, b * s
        g = out_grad.contiguous()
        if g._code != x2._code:
            g = g.astype(x2._dtype)
        hv = (b, heads, s, dh), (s * H, dh, H, 1)
        q, k, v = (qkv._view(hv[0], hv[1], i * rows * H) for i in range(3))
        go = g._view(hv[0], hv[1])
        dqkv = CudaTensor._new((3, rows, H), x2._dtype)
        dq, dk, dv = (dqkv._view(hv[0], hv[1], i * rows * H) for i in range(3))
        if isinstance(probs, tuple):
            # fused path: dQ, dK, dV from one residency of Q, K, V, dO (probabilities recomputed from the saved LSE)
            out, lse = probs
            bgs = [_direct_grad(bias, x2._code) for bias in (bq, bk, bv)]
            in_kernel = all(t is not None for t in bgs) and 'bias_epi' not in _DISABLED
            if in_kernel:
                for t in bgs:
                    rt.side_join_if_written(t)
            rt.api.attention_bwd(x2._code, qkv.ptr, out.ptr, g.ptr, lse.ptr, b, s, heads, dh, scale, dqkv.ptr,
                                 *([t.ptr for t in bgs] if in_kernel else [None, None, None]))
            return self_attention._projection_backward(ctx, dqkv, x2, rows, H, xshape, bias_done=in_kernel)
        _gemm(_swap_last(probs), go, out=dv, cls='A')                   # dV = P^T dO
        ds = CudaTensor._new(probs._shape, x2._dtype)
        if not _attention_gemm(go, _swap_last(v), ds, 4, scale, aux=probs):   # dS straight from the dP GEMM's epilogue
            dp = _gemm(go, _swap_last(v), cls='A')                      # dP = dO V^T
            rt.api.softmax_bwd(x2._code, probs.ptr, dp.ptr, ds.ptr, probs._numel // s, s, scale)
            del dp
        _gemm(ds, k, out=dq, cls='A')                                   # dQ = dS K
        _gemm(_swap_last(ds), q, out=dk, cls='A')                       # dK = dS^T Q
        return self_attention._projection_backward(ctx, dqkv, x2, rows, H, xshape)

    @staticmethod
    def _projection_backward(ctx, dqkv, x2, rows, H, xshape, bias_done=False):
        """dX, dW_g, db_g of the three projections from the stacked (3, rows, H) gradient of Q, K, V
        (``bias_done``: the attention kernel already added the column sums to the bias gradients)."""
        wq, bq, wk, bk, wv, bv = ctx._parents[1:7]
        parts = [dqkv._view((rows, H), (H, 1), i * rows * H) for i in range(3)]
        ws = (wq, wk, wv)
        xin = ctx._parents[0]
        xg = _direct_grad(xin, x2._code) if isinstance(xin, CudaTensor) and xin._shape == tuple(xshape) else None
        if xg is not None:
            # dX = sum_g dY_g W_g added into the gradient x already received through the residual branch
            rt.side_join_if_written(xg)
            _gemm_grouped(parts, list(ws), [_fold_rows(xg)] * 3, accumulate=True, cls='X')
            dx = Function.ACCUMULATED
        else:
            dx = CudaTensor._new((rows, H), x2._dtype)
            _gemm_grouped(parts, list(ws), [dx] * 3, cls='X')           # dX = sum_g dY_g W_g
            dx = _with_shape(dx, xshape)
        wgs = [_direct_grad(w, x2._code) for w in ws]
        bgs = [_direct_grad(bias, x2._code) for bias in (bq, bk, bv)]
        pt = [_swap_last(t) for t in parts]
        if all(w is not None for w in wgs) and all(b is not None for b in bgs) and H > 1:
            # dW_g += dY_g^T X and db_g += colsum(dY_g), straight into the arena, on the side stream
            with rt.side_stream(dqkv, x2, writes=tuple(wgs) + tuple(bgs)):
                _gemm_grouped(pt, [x2] * 3, wgs, accumulate=True, cls='W')
                if not bias_done:
                    for part, bg in zip(parts, bgs):
                        _bias_grad(part._code, part.ptr, bg.ptr, rows, H, H)
            return (dx,) + (Function.ACCUMULATED,) * 6
        if all(w is not None for w in wgs):
            _gemm_grouped(pt, [x2] * 3, wgs, accumulate=True, cls='W')
            dws = [Function.ACCUMULATED] * 3
        else:
            dws = [CudaTensor._new((H, H), x2._dtype) for _ in range(3)]
            _gemm_grouped(pt, [x2] * 3, dws, cls='W')
        dbs = []
        for part, bg in zip(parts, bgs):
            if bias_done:
                dbs.append(Function.ACCUMULATED)
            elif bg is not None and H > 1:
                _bias_grad(part._code, part.ptr, bg.ptr, rows, H, H)
                dbs.append(Function.ACCUMULATED)
            else:
                dbs.append(_reduce(RED['SUM'], part, (0,), False))
        return dx, dws[0], dbs[0], dws[1], dbs[1], dws[2], dbs[2]


# ---------------------------------------------------------------------------------------------------
# indexing (cpu/ops.py:234-255, opencl/ops.py:292-331)
def _is_index_array(v):
    return isinstance(v, (CudaTensor, np.ndarray, list, range)) and not isinstance(v, (str, bytes))


def _parse_index(shape, idx):
    """Split a numpy-style index into a basic part (view) and integer-array parts.

    Returns (basic, arrays): ``basic`` is a list with one entry per source dim -- int, slice or None
    (None marks a dim consumed by an index array) -- and ``arrays`` is [(dim, array_like)].
    """
    if not isinstance(idx, tuple):
        idx = (idx,)
    nd = len(shape)
    n_specified = builtins_sum(1 for i in idx if i is not Ellipsis and i is not None)
    out, dim = [], 0
    for i in idx:
        if i is Ellipsis:
            for _ in range(nd - n_specified):
                out.append(slice(None))
                dim += 1
        elif i is None:
            raise IndexError("newaxis is not supported in CudaTensor indexing")
        else:
            out.append(i)
            dim += 1
    while len(out) < nd:
        out.append(slice(None))
    if len(out) > nd:
        raise IndexError("too many indices for tensor of shape %s" % (shape,))
    arrays = [(d, v) for d, v in enumerate(out) if _is_index_array(v)]
    if arrays:
        # integer scalars mixed with index arrays broadcast with them, as in numpy
        arrays = [(d, v) for d, v in enumerate(out) if _is_index_array(v) or isinstance(v, (int, np.integer))]
    return out, arrays


import builtins as _bi  # noqa: E402
builtins_sum = _bi.sum


def _basic_view(t, basic):
    """Apply ints and slices (entries that are None are left untouched)."""
    shape, strides, offset = [], [], t._offset
    for d, i in enumerate(basic):
        n, st = t._shape[d], t._strides[d]
        if i is None:
            shape.append(n)
            strides.append(st)
        elif isinstance(i, (int, np.integer)):
            i = int(i)
            if i < -n or i >= n:
                raise IndexError("index %d is out of bounds for axis %d with size %d" % (i, d, n))
            offset += (i % n) * st
        elif isinstance(i, slice):
            lo, hi, step = i.indices(n)
            cnt = len(range(lo, hi, step))
            shape.append(cnt)
            strides.append(st * step)
            offset += lo * st if cnt else 0
        else:
            raise IndexError("unsupported index %r" % (i,))
    return t._view(tuple(shape), tuple(strides), offset)


def _index_plan(t, idx):
    """Reduce a fancy index to (source rows view, row numbers) -- see _gather."""
    basic, arrays = _parse_index(t._shape, idx)
    if not arrays:
        return _basic_view(t, basic), None
    idx_dev = list(basic)
    dims = [d for d, _ in arrays]
    if dims != list(range(dims[0], dims[0] + len(dims))):
        raise IndexError("index arrays must address adjacent dims on the cuda backend")
    if dims[0] != 0:
        raise IndexError("index arrays must start at dim 0 on the cuda backend (got dim %d)" % dims[0])
    for d, _ in arrays:
        basic[d] = None
    src = _basic_view(t, basic)
    k = len(arrays)
    # the trailing dims must be dense rows
    tail_shape = src._shape[k:]
    tail_ok = list(src._strides[k:]) == list(contiguous_strides(tail_shape))
    if not tail_ok:
        src = src.contiguous()
    row_len = _prod(tail_shape)
    # index arrays -> device int tensors broadcast to one shape
    dev, bshape = [], ()
    for d, v in arrays:
        if isinstance(v, CudaTensor):
            iv = v
        else:
            a = np.asarray(list(v) if isinstance(v, range) else v)
            if a.dtype == np.bool_:
                raise IndexError("boolean masks are not supported on the cuda backend")
            if a.dtype.kind not in 'iu':
                raise IndexError("index arrays must be integers")
            iv = CudaTensor.from_numpy(a.astype(np.int64) if a.dtype.kind == 'u' and a.dtype.itemsize > 1 else a,
                                       requires_grad=False)
        if iv._code not in (rt.I32, rt.I64, rt.I16, rt.U8, rt.I8):
            raise IndexError("index tensors must be integers, got %s" % iv._dtype)
        dev.append(iv)
        idx_dev[d] = iv
        bshape = _bshape(bshape, iv._shape)
    n_idx = _prod(bshape)
    if k == 1 and dev[0]._contig and dev[0]._shape == bshape:
        rows, row_stride, n_rows = dev[0], src._strides[0], src._shape[0]
    else:
        # fold several index arrays (or broadcast ones) into element offsets on the device
        mats = []
        for iv in dev:
            if iv._shape != bshape or not iv._contig:
                full = CudaTensor._new(bshape, iv._dtype)
                rt.api.cast(iv._code, iv._code, len(bshape), i64arr(bshape), iv.ptr, i64arr(_bstrides(iv, bshape)),
                            full.ptr, None)
                iv = full
            mats.append(iv)
        lin = CudaTensor._new(bshape, np.int64)
        ptrs = (C.c_void_p * k)(*[m.ptr for m in mats])
        dts = (C.c_int * k)(*[m._code for m in mats])
        rt.api.index_linearize(k, ptrs, dts, i64arr(src._shape[:k]), i64arr(src._strides[:k]), n_idx, lin.ptr)
        rows, row_stride, n_rows = lin, 1, 1 << 62
        dev = mats + [lin]
    return src, (rows, row_stride, n_rows, n_idx, row_len, bshape, tail_shape, tuple(idx_dev))


@CudaTensor.register_op("__getitem__")
class getitem(Function):
    def forward(ctx, a, idx):
        src, plan = _index_plan(a, idx)
        if plan is None:
            ctx.save_for_backward(a._shape, a._dtype, idx)
            return src
        rows, row_stride, n_rows, n_idx, row_len, bshape, tail_shape, idx_dev = plan
        # keep the uploaded index arrays so backward does not stage them again
        ctx.save_for_backward(a._shape, a._dtype, idx_dev)
        ctx.source = a
        out = CudaTensor._new(bshape + tail_shape, a._dtype)
        rt.api.gather_rows(a._code, rows._code, src.ptr, n_rows, row_stride, rows.ptr, n_idx, row_len, out.ptr)
        return out

    def backward(ctx, out_grad):
        shape, dtype, idx = ctx.get_saved_tensors()
        a = getattr(ctx, 'source', None)
        if a is not None and a.requires_grad and a.grad is not None and a.grad._contig and a.grad._shape == shape \
                and a.grad._code == out_grad._code and a.grad._code in (rt.F32, rt.F64):
            # scatter-add straight into the existing gradient of the table (skips zero-filling and adding a
            # table-sized temporary: 94 MB for BERT's word embeddings)
            rt.side_join_if_written(a.grad)     # e.g. a table tied to a Linear weight whose dW runs on the side stream
            src, plan = _index_plan(a.grad, idx)
            rows, row_stride, n_rows, n_idx, row_len, bshape, tail_shape, _keep = plan
            if src._data is a.grad._data:
                g = out_grad.contiguous()
                rt.api.scatter_add_rows(g._code, rows._code, src.ptr, n_rows, row_stride, rows.ptr, n_idx, row_len, g.ptr)
                return Function.ACCUMULATED
        grad = CudaTensor.zeros(shape, dtype=np.float32 if dtype.kind != 'f' else dtype, requires_grad=False)
        src, plan = _index_plan(grad, idx)
        g = out_grad if out_grad._code == grad._code else out_grad.astype(grad._dtype)
        if plan is None:
            # basic index: the view does not overlap itself, assignment == accumulation
            _assign(src, g)
        else:
            rows, row_stride, n_rows, n_idx, row_len, bshape, tail_shape, _keep = plan
            if src._data is not grad._data:
                raise IndexError("scatter-add into a non-dense slice is not supported")
            g = g.contiguous()
            rt.api.scatter_add_rows(grad._code, rows._code, src.ptr, n_rows, row_stride, rows.ptr, n_idx, row_len,
                                    g.ptr)
        grad._temp = True
        return grad


@CudaTensor.register_op("__setitem__")
class setitem(Function):
    def forward(ctx, a, idx, val):
        _written(a)
        src, plan = _index_plan(a, idx)
        if plan is None:
            if _is_scalar(val):
                src._fill_value(val)
            else:
                _assign(src, _as_tensor(val, a))
            return a._view(a._shape, a._strides)
        rows, row_stride, n_rows, n_idx, row_len, bshape, tail_shape, _keep = plan
        if src._data is not a._data:
            raise IndexError("fancy assignment into a non-dense slice is not supported")
        if _is_scalar(val):
            rt.api.scatter_set_rows(a._code, rows._code, src.ptr, n_rows, row_stride, rows.ptr, n_idx, row_len,
                                    None, float(val))
        else:
            val = _as_tensor(val, a)
            if val._code != a._code:
                val = val.astype(a._dtype)
            full = bshape + tail_shape
            if val._shape != full or not val._contig:
                dense = CudaTensor._new(full, a._dtype)
                rt.api.cast(a._code, a._code, len(full), i64arr(full), val.ptr, i64arr(_bstrides(val, full)),
                            dense.ptr, None)
                val = dense
            rt.api.scatter_set_rows(a._code, rows._code, src.ptr, n_rows, row_stride, rows.ptr, n_idx, row_len,
                                    val.ptr, 0.0)
        return a._view(a._shape, a._strides)


# ---------------------------------------------------------------------------------------------------
# fused layers for the BERT path (new op names; nn/loss use them when the tensor class has them)
@CudaTensor.register_op(overwrite=True)
class softmax(Function):
    """exp(t - max) / sum along ``axis`` in one kernel (generic form: ops.py:62-66 of the reference).

    ``scale`` multiplies the input first (fuses attention's 1/sqrt(d), examples/bert.py:79).
    """
    def forward(ctx, t, axis=-1, scale=1.0):
        t = _float_like(t)
        nd = len(t._shape)
        axis = axis % nd
        ctx.axis, ctx.scale, ctx.nd = axis, float(scale), nd
        x = t if axis == nd - 1 else t.transpose(*[i for i in range(nd) if i != axis], axis)
        x = x.contiguous()
        cols = x._shape[-1]
        y = CudaTensor._new(x._shape, x._dtype)
        if x._numel:
            rt.api.softmax_fwd(x._code, x.ptr, y.ptr, x._numel // cols, cols, float(scale))
        y._temp = False
        ctx.save_for_backward(y)
        if axis != nd - 1:
            inv = list(range(axis)) + [nd - 1] + list(range(axis, nd - 1))
            return y.transpose(*inv)
        return y._view(y._shape, y._strides)

    def backward(ctx, out_grad):
        y, = ctx.get_saved_tensors()
        axis, nd = ctx.axis, ctx.nd
        g = out_grad if axis == nd - 1 else out_grad.transpose(*[i for i in range(nd) if i != axis], axis)
        g = g.contiguous()
        if g._code != y._code:
            g = g.astype(y._dtype)
        cols = y._shape[-1]
        dx = CudaTensor._new(y._shape, y._dtype)
        if y._numel:
            rt.api.softmax_bwd(y._code, y.ptr, g.ptr, dx.ptr, y._numel // cols, cols, ctx.scale)
        if axis != nd - 1:
            inv = list(range(axis)) + [nd - 1] + list(range(axis, nd - 1))
            return dx.transpose(*inv)
        return dx


@CudaTensor.register_op()
class layernorm(Function):
    """(x - mean) / sqrt(var + eps) * weight + bias over the last axis (nn.py:109-124 of the reference)."""
    def forward(ctx, x, weight, bias, eps=1e-5):
        x = _float_like(x).contiguous()
        cols = x._shape[-1]
        assert weight._shape == (cols,) and bias._shape == (cols,)
        rows = x._numel // cols
        y = CudaTensor._new(x._shape, x._dtype)
        mean, rstd = CudaTensor._new((rows,), x._dtype), CudaTensor._new((rows,), x._dtype)
        w, b = weight.contiguous(), bias.contiguous()
        rt.api.layernorm_fwd(x._code, x.ptr, w.ptr, b.ptr, y.ptr, mean.ptr, rstd.ptr, rows, cols, float(eps))
        x._mark_shared()
        ctx.save_for_backward(x, w, mean, rstd)
        return y

    def backward(ctx, out_grad):
        x, w, mean, rstd = ctx.get_saved_tensors()
        return _ln_backward(x, w, mean, rstd, ctx._parents[1], ctx._parents[2], out_grad)


def _ln_backward(x, w, mean, rstd, weight, bias, out_grad):
    """(dx, d(gamma), d(beta)) of a layer norm over the last axis; the parameter gradients are added straight
    into existing gradient buffers when there are any."""
    g = out_grad.contiguous()
    cols = x._shape[-1]
    rows = x._numel // cols
    dx = CudaTensor._new(x._shape, x._dtype)
    # column sums of dx, formed by the same kernel: if x came out of a Linear layer (the residual blocks) they are that
    # layer's bias gradient, and its backward picks them up from the buffer's note instead of re-reading dx
    cs = None
    if x._code == rt.F32 and cols % 4 == 0 and 8 <= cols <= 1024 and rows >= 64 and 'bias_epi' not in _DISABLED:
        cs = CudaTensor._new((cols,), x._dtype, requires_grad=False)
        _root(dx)._colsum = (cs, dx.ptr, rows, cols, cols)
    cs_ptr = cs.ptr if cs is not None else None
    wg, bg = weight.grad, bias.grad
    if weight.requires_grad and bias.requires_grad and wg is not None and bg is not None and wg._contig \
            and bg._contig and wg._code == bg._code == x._code and wg._shape == bg._shape == (cols,):
        # d(gamma), d(beta) are added straight into the existing gradients
        rt.api.layernorm_bwd(x._code, x.ptr, w.ptr, mean.ptr, rstd.ptr, g.ptr, dx.ptr, wg.ptr, bg.ptr, rows, cols, 1,
                             cs_ptr)
        # (the library sums the per-CTA partials into wg / bg on the side stream)
        rt._side_dirty.add(wg.ptr)
        rt._side_dirty.add(bg.ptr)
        return dx, Function.ACCUMULATED, Function.ACCUMULATED
    dw, db = CudaTensor._new((cols,), x._dtype), CudaTensor._new((cols,), x._dtype)
    rt.api.layernorm_bwd(x._code, x.ptr, w.ptr, mean.ptr, rstd.ptr, g.ptr, dx.ptr, dw.ptr, db.ptr, rows, cols, 0, cs_ptr)
    return dx, dw, db


@CudaTensor.register_op()
class add_layernorm(Function):
    """layernorm(a + b) over the last axis: the residual connections of BertAttention / BertLayer
    (examples/bert.py:113,158 of the reference: LayerNorm(hidden + hidden_in)) with the add folded into the
    normalisation kernel (one pass writes the sum, which backward needs, and the normalised result)."""
    def forward(ctx, a, b, weight, bias, eps=1e-5):
        a, b = _float_like(a), _float_like(b)
        if a._shape != b._shape or a._code != b._code:
            raise ValueError("add_layernorm: operands must share one shape and dtype (%s %s / %s %s)"
                             % (a._shape, a._dtype, b._shape, b._dtype))
        a, b = a.contiguous(), b.contiguous()
        cols = a._shape[-1]
        assert weight._shape == (cols,) and bias._shape == (cols,)
        rows = a._numel // cols
        s = CudaTensor._new(a._shape, a._dtype)
        y = CudaTensor._new(a._shape, a._dtype)
        mean, rstd = CudaTensor._new((rows,), a._dtype), CudaTensor._new((rows,), a._dtype)
        w, bb = weight.contiguous(), bias.contiguous()
        rt.api.add_layernorm_fwd(a._code, a.ptr, b.ptr, s.ptr, w.ptr, bb.ptr, y.ptr, mean.ptr, rstd.ptr, rows, cols,
                                 float(eps))
        s._temp = False
        ctx.save_for_backward(s, w, mean, rstd)
        return y

    def backward(ctx, out_grad):
        s, w, mean, rstd = ctx.get_saved_tensors()
        dx, dw, db = _ln_backward(s, w, mean, rstd, ctx._parents[2], ctx._parents[3], out_grad)
        return dx, dx, dw, db       # the sum's gradient goes to both addends


def cross_entropy_forward(logits, labels):
    """Fused log-softmax + NLL over the last axis.  Returns (mean loss tensor of shape (), saved state)."""
    x = _float_like(logits)
    assert len(x._shape) == 2, "fused cross entropy expects (rows, classes) logits"
    if x._strides[1] != 1 or x._strides[0] < x._shape[1]:
        x = x.contiguous()
    rows, cols = x._shape
    lab = labels if isinstance(labels, CudaTensor) else CudaTensor.from_numpy(np.asarray(labels), requires_grad=False)
    lab = lab.contiguous()
    if lab._code not in (rt.I32, rt.I64, rt.I16):
        lab = lab.astype(np.int64)
    loss_rows, lse = CudaTensor._new((rows,), x._dtype), CudaTensor._new((rows,), x._dtype)
    rt.api.cross_entropy_fwd(x._code, lab._code, x.ptr, x._strides[0], lab.ptr, loss_rows.ptr, lse.ptr, rows, cols)
    loss = _reduce(RED['SUM'], loss_rows, (0,), False, scale=1.0 / rows)
    x._mark_shared()
    return loss, (x, lab, lse)


def cross_entropy_backward(saved, out_grad):
    # (forming this gradient inside the forward kernel, while each row is cache resident, was measured slower
    #  than the two separate sweeps: 8.63 vs 8.55 ms per BERT step -- the forward grid has to shrink to keep
    #  the rows in flight inside L2)
    x, lab, lse = saved
    rows, cols = x._shape
    g = out_grad.contiguous()
    if g._code != x._code:
        g = g.astype(x._dtype)
    # same (possibly padded) row pitch as the logits, so the backward GEMMs can read it through TMA
    ld = x._strides[0]
    full = CudaTensor._new((rows, ld), x._dtype)
    dx = full._view((rows, cols), (ld, 1))
    dx._temp = True
    rt.api.cross_entropy_bwd(x._code, lab._code, x.ptr, ld, lab.ptr, lse.ptr, g.ptr, dx.ptr, ld, rows, cols)
    return dx


CudaTensor.fused_cross_entropy = staticmethod(cross_entropy_forward)
CudaTensor.fused_cross_entropy_backward = staticmethod(cross_entropy_backward)


# ---------------------------------------------------------------------------------------------------
# convolution (cpu/ops.py:298-356 of the reference): unfold (lg_im2col) + matmul; backward = two matmuls + fold
# (lg_col2im, a gather: one launch instead of one strided add per kernel offset)
def _conv_geometry(t_shape, k_shape, strides):
    """(leading extent, window input dims, positions per window dim) of ``n = len(k_shape)`` trailing dims."""
    n = len(k_shape)
    in_dims = tuple(t_shape[-n:])
    for d, k, s in zip(in_dims, k_shape, strides):
        if k > d or s < 1:
            raise ValueError("conv: kernel %s with strides %s does not fit an input of %s" % (k_shape, strides, in_dims))
    pos = tuple((d - k) // s + 1 for d, k, s in zip(in_dims, k_shape, strides))
    return _prod(t_shape[:-n]), in_dims, pos


@CudaTensor.register_op()
class conv(Function):
    def forward(ctx, t, kernel, strides=1):
        """``t``: (..., C, *spatial); ``kernel``: (out_channels, C, *k).  The channel dim is a window dim with kernel
        extent C and stride 1, exactly as in the reference, so any number of spatial dims up to 3 works."""
        t, kernel = _promote(t, kernel)
        n = len(kernel._shape) - 1
        if isinstance(strides, (int, np.integer)):
            strides = (int(strides),) * n
        elif len(strides) == n - 1:
            strides = (1,) + tuple(int(s) for s in strides)
        else:
            strides = tuple(int(s) for s in strides)
        if not (len(t._shape) >= n == len(strides)) or n > 4:
            raise ValueError("conv: input %s, kernel %s, strides %s do not go together" % (t._shape, kernel._shape, strides))
        k_shape = kernel._shape[1:]
        lead, in_dims, pos = _conv_geometry(t._shape, k_shape, strides)
        n_pos, n_k, out_c = _prod(pos), _prod(k_shape), kernel._shape[0]
        x = t.contiguous()
        cols = CudaTensor._new((lead * n_pos, n_k), x._dtype)
        if cols._numel:
            rt.api.im2col(x._code, n, lead, i64arr(in_dims), i64arr(k_shape), i64arr(strides), x.ptr, cols.ptr)
        flat_w = kernel.contiguous()._view((out_c, n_k), None)
        y = _gemm(cols, _swap_last(flat_w))                  # (lead * positions, out_channels)
        cols._mark_shared()
        ctx.save_for_backward(cols, flat_w, t._shape, kernel._shape, strides, (lead, in_dims, pos))
        # (..., 1, *spatial positions, out_channels) -> channels take the place of the collapsed channel dim
        full = tuple(t._shape[:-n]) + pos + (out_c,)
        y = _with_shape(y, full)
        nd = len(full)
        perm = list(range(nd))
        perm[nd - n - 1], perm[nd - 1] = perm[nd - 1], perm[nd - n - 1]
        y = y.transpose(*perm)
        assert y._shape[-1] == 1
        return y._view(y._shape[:-1], y._strides[:-1])

    def backward(ctx, out_grad):
        cols, flat_w, in_shape, w_shape, strides, (lead, in_dims, pos) = ctx.get_saved_tensors()
        n = len(w_shape) - 1
        g = out_grad if out_grad._code == cols._code else out_grad.astype(cols._dtype)
        nd = len(g._shape)
        # channels last again: (..., *positions, out_channels) -> (lead * positions, out_channels)
        perm = [i for i in range(nd) if i != nd - n] + [nd - n]
        flat_g = g.transpose(*perm).contiguous()._view((lead * _prod(pos), w_shape[0]), None)
        dcols = _gemm(flat_g, flat_w, cls='X')               # gradient of the unfolded input
        w_grad = _with_shape(_gemm(_swap_last(flat_g), cols, cls='W'), w_shape)
        x_grad = CudaTensor._new(in_shape, cols._dtype, requires_grad=False)
        if x_grad._numel:
            rt.api.col2im(cols._code, n, lead, i64arr(in_dims), i64arr(w_shape[1:]), i64arr(strides), dcols.ptr, x_grad.ptr)
        return x_grad, w_grad


# measurement switch (see _DISABLED above): drop fused nodes so that models fall back to the composed path
for _n, _attr in (('add_ln', 'add_layernorm'), ('mlp', 'mlp_gelu'), ('attn', 'self_attention')):
    if _n in _DISABLED and hasattr(CudaTensor, _attr):
        delattr(CudaTensor, _attr)
