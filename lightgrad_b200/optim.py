"""Optimizers with the reference's update rules (lightgrad/optim.py:3-52).

Two execution paths with identical arithmetic:

* generic  -- the reference's per-parameter loop ``p += compute_delta(p.grad, i)`` built from tensor
  operators (works for any backend);
* fused    -- when every parameter is a float32 CudaTensor the parameters, their gradients and the
  optimizer state are moved into flat device arenas at construction (the parameter objects keep
  their identity and become views into the arena) and one kernel updates everything per step.
  The flat gradient arena is also what the data-parallel wrapper all-reduces.

Kept on purpose: Adam's step counter ``t`` advances once per PARAMETER per step
(optim.py:36-37), so parameter i of P sees t = (step-1)*P + i + 1 in its bias correction.
"""
import numpy as np
from .autograd import Gradients, AbstractTensor

_ALIGN = 64  # elements; keeps every tensor of an arena 256-byte aligned


class _Arena(object):
    """Flat float32 device arenas for parameters and gradients (cuda backend only)."""

    def __init__(self, params):
        from .autograd.cuda import runtime as rt
        from .autograd.cuda.tensor import CudaTensor
        self.rt, self.T = rt, CudaTensor
        self.offsets, off = [], 0
        for p in params:
            self.offsets.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.total = off
        self.param_buf = rt.Buffer(max(off, 1) * 4)
        self.grad_buf = rt.Buffer(max(off, 1) * 4)
        rt.api.memset(self.param_buf.ptr, 0, off * 4)
        rt.api.memset(self.grad_buf.ptr, 0, off * 4)
        for p, o in zip(params, self.offsets):
            n = p.numel()
            src = p.contiguous()
            rt.api.memcpy_d2d(self.param_buf.ptr + o * 4, src.ptr, n * 4)
            old_grad = p.grad
            # rebind the parameter object onto its arena window (identity preserved)
            p._set_data(rt.ArenaSlice(self.param_buf, o * 4, n * 4))
            p._offset, p._strides, p._contig = 0, _contig_strides(p.shape), True
            g = CudaTensor(rt.ArenaSlice(self.grad_buf, o * 4, n * 4), p.shape, None, 0, np.float32, False)
            if old_grad is not None:
                rt.api.memcpy_d2d(g.ptr, old_grad.contiguous().ptr, n * 4)
            p._grad = g
        self.seg_end = None

    def rebase(self, params, param_ptr, grad_ptr, keep=None):
        """Move both arenas to memory the caller provides (the NVLink multicast region of the data-parallel
        wrapper): contents are copied, the parameter / gradient objects keep their identity and become views of the
        new arenas.  Captured CUDA graphs that recorded the old addresses must be re-captured."""
        rt = self.rt
        new_p = rt.ExternalBuffer(param_ptr, max(self.total, 1) * 4, keep)
        new_g = rt.ExternalBuffer(grad_ptr, max(self.total, 1) * 4, keep)
        self.adopt_grads(params)
        rt.api.memcpy_d2d(new_p.ptr, self.param_buf.ptr, self.total * 4)
        rt.api.memcpy_d2d(new_g.ptr, self.grad_buf.ptr, self.total * 4)
        rt.synchronize()
        self.param_buf, self.grad_buf = new_p, new_g
        for p, o in zip(params, self.offsets):
            n = p.numel()
            p._set_data(rt.ArenaSlice(new_p, o * 4, n * 4))
            p._offset = 0
            p._grad = self.T(rt.ArenaSlice(new_g, o * 4, n * 4), p.shape, None, 0, np.float32, False)

    def state(self):
        b = self.rt.Buffer(max(self.total, 1) * 4)
        self.rt.api.memset(b.ptr, 0, self.total * 4)
        return b

    def adopt_grads(self, params):
        """Make sure every p.grad still aliases the gradient arena (it does unless user code replaced it)."""
        rt = self.rt
        for p, o in zip(params, self.offsets):
            want = self.grad_buf.ptr + o * 4
            g = p.grad
            if g is None:
                p._grad = self.T(rt.ArenaSlice(self.grad_buf, o * 4, p.numel() * 4), p.shape, None, 0, np.float32, False)
                rt.api.memset(want, 0, p.numel() * 4)
            elif g.ptr != want or not g.is_contiguous():
                rt.api.memcpy_d2d(want, g.contiguous().ptr, p.numel() * 4)
                p._grad = self.T(rt.ArenaSlice(self.grad_buf, o * 4, p.numel() * 4), p.shape, None, 0, np.float32, False)

    def segments(self, params):
        if self.seg_end is None:
            ends = np.array([o + p.numel() for p, o in zip(params, self.offsets)], dtype=np.int64)
            # padding between tensors belongs to the preceding tensor (values there are never read back)
            ends[:-1] = np.array(self.offsets[1:], dtype=np.int64)
            ends[-1] = self.total
            self.seg_end = self.T.from_numpy(ends, requires_grad=False)
        return self.seg_end


def _contig_strides(shape):
    st, acc = [], 1
    for s in reversed(shape):
        st.append(acc)
        acc *= s
    return tuple(reversed(st))


def _can_fuse(params):
    try:
        from .autograd.cuda.tensor import CudaTensor
    except Exception:
        return False
    return len(params) > 0 and all(isinstance(p, CudaTensor) and p.dtype == np.float32 and p.requires_grad
                                   for p in params)


class Optimizer(object):

    def __init__(self, parameters, fused=None):
        self.parameters = tuple(parameters)
        assert all(isinstance(p, AbstractTensor) for p in self.parameters)
        want = _can_fuse(self.parameters) if fused is None else fused
        if want and len(set(id(p) for p in self.parameters)) != len(self.parameters):
            # a tensor listed twice (tied weights) would get two arena slots, one of them orphaned.  The
            # reference's loop updates such a tensor twice per step, each time with its own m / v / t
            # (optim.py:10-13); only the generic per-parameter path reproduces that, so it is used here.
            if fused:
                raise ValueError("fused optimizer: every parameter must be listed once (tied weights: use fused=False)")
            want = False
        self.arena = _Arena(self.parameters) if want else None

    def zero_grad(self):
        if self.arena is not None:
            a = self.arena
            a.adopt_grads(self.parameters)
            a.rt.api.memset(a.grad_buf.ptr, 0, a.total * 4)
            return
        for p in self.parameters:
            p.zero_grad()

    def step(self):
        Gradients.disable()
        try:
            if self.arena is not None:
                self.arena.adopt_grads(self.parameters)
                self._fused_step(self.arena)
                self.arena.param_buf._bf16 = None      # a bf16 staging copy of the parameters is stale now
            else:
                for i, p in enumerate(self.parameters):
                    p += self.compute_delta(p.grad, i)
        finally:
            Gradients.enable()

    def compute_delta(self, grad, idx):
        raise NotImplementedError()

    def _fused_step(self, arena):
        raise NotImplementedError()


class SGD(Optimizer):
    """Stochastic Gradient Descent (optional momentum)."""

    def __init__(self, parameters, lr, momentum=0.0, fused=None):
        Optimizer.__init__(self, parameters, fused)
        self.prev_deltas = [0] * len(self.parameters)
        self.lr, self.momentum = lr, momentum
        self._delta = None

    def compute_delta(self, grad, i):
        self.prev_deltas[i] = -self.lr * grad + self.momentum * self.prev_deltas[i]
        return self.prev_deltas[i]

    def _fused_step(self, a):
        if self.momentum != 0.0 and self._delta is None:
            self._delta = a.state()
        a.rt.api.sgd_step(a.param_buf.ptr, a.grad_buf.ptr, self._delta.ptr if self._delta is not None else None,
                          a.total, float(self.lr), float(self.momentum))

    def _bucket_step_range(self, a, i0, i1, lo, hi, last):
        """Single GPU: SGD update of arena elements [lo, hi) by the small-footprint kernel on the collective stream, so
        that it runs beside the rest of backward (lg_bucket_step)."""
        if self.momentum != 0.0 and self._delta is None:
            self._delta = a.state()
        a.rt.api.bucket_step(2, a.param_buf.ptr, a.grad_buf.ptr, self._delta.ptr if self._delta is not None else None,
                             None, lo, hi, 0, None, None, float(self.lr), 0.0, 0.0, 0.0, float(self.momentum), i0, 0)

    def _mc_exchange_range(self, a, region, i0, i1, lo, hi, rank, world, last):
        """Parameters i0 .. i1-1 = arena elements [lo, hi): gradient reduce-scatter, SGD update of this rank's share
        and parameter all-gather in one kernel over the NVLink multicast region (lg_mc_exchange_step)."""
        if self.momentum != 0.0 and self._delta is None:
            self._delta = a.state()
        goff, poff, foff = region
        a.rt.api.mc_exchange_step(2, goff, poff, foff, lo, hi, rank, world,
                                  self._delta.ptr if self._delta is not None else None, None, 0, None, None,
                                  float(self.lr), 0.0, 0.0, 0.0, float(self.momentum), i0, 0)


class Adam(Optimizer):
    """ADAptive Moment estimation."""
    _belief = 0

    def __init__(self, parameters, lr, beta1=0.9, beta2=0.999, eps=1e-8, fused=None):
        Optimizer.__init__(self, parameters, fused)
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.t = 0
        self.m = [0] * len(self.parameters)
        self.v = [0] * len(self.parameters)
        self._m = self._v = None
        if self.arena is not None:
            self._init_state(self.arena)

    def _init_state(self, a):
        # allocated and zeroed up front, on the compute stream: a per-bucket update issued on the collective stream
        # (DataParallel.backward_and_step) must never be the one that creates -- and memsets -- this state
        self._m, self._v = a.state(), a.state()
        # the step counter lives on the device (and is advanced there) so that a step captured into
        # a CUDA graph keeps counting when it is replayed
        self._t_dev = a.T.from_numpy(np.array([self.t], dtype=np.int64), requires_grad=False)
        a.segments(self.parameters)          # uploaded now: the first step may already be under CUDA-graph capture

    def steps_taken(self):
        """Parameter updates performed so far (``t`` of the reference), read from the device counter when the
        fused path keeps it there -- the host-side ``self.t`` does not advance while a captured step is replayed."""
        if self._m is not None:
            return int(self._t_dev.numpy()[0])
        return self.t

    def compute_delta(self, grad, i):
        self.t += 1
        self.m[i] = self.b1 * self.m[i] + (1 - self.b1) * grad
        self.v[i] = self.b2 * self.v[i] + (1 - self.b2) * self._second_moment_input(grad, i) ** 2
        m, v = self.m[i] / (1 - self.b1 ** self.t), self.v[i] / (1 - self.b2 ** self.t)
        return -self.lr * m / (v ** 0.5 + self.eps)

    def _second_moment_input(self, grad, i):
        return grad

    def _fused_step(self, a):
        self._fused_step_range(a, 0, len(self.parameters), last=True)

    def _fused_step_range(self, a, i0, i1, last):
        """Update parameters i0 .. i1-1 (a contiguous window of the arenas).  The calls of one step may come in any
        order; the one with ``last=True`` (issued last) advances the step counter for the whole step."""
        P = len(self.parameters)
        if self._m is None:
            self._init_state(a)
        seg = a.segments(self.parameters)
        lo = a.offsets[i0]
        hi = a.offsets[i1] if i1 < P else a.total
        # parameter i uses t = t_dev + i + 1 (the reference bumps t once per parameter); the kernel
        # derives the two bias corrections from t itself, so nothing is uploaded per step
        a.rt.api.adam_step(self._belief, a.param_buf.ptr + lo * 4, a.grad_buf.ptr + lo * 4, self._m.ptr + lo * 4,
                           self._v.ptr + lo * 4, hi - lo, i1 - i0, seg.ptr + i0 * 8, self._t_dev.ptr, float(self.lr),
                           float(self.b1), float(self.b2), float(self.eps), lo, i0, P if last else 0)
        if last:
            self.t += P

    def _bucket_step_range(self, a, i0, i1, lo, hi, last):
        """Single GPU: Adam / AdaBelief on arena elements [lo, hi) (parameters i0 .. i1-1) by the small-footprint kernel
        on the collective stream -- beside the rest of backward instead of after it (lg_bucket_step).  Step counter
        and bias corrections as in ``_fused_step_range``."""
        P = len(self.parameters)
        if self._m is None:
            self._init_state(a)
        seg = a.segments(self.parameters)
        a.rt.api.bucket_step(self._belief, a.param_buf.ptr, a.grad_buf.ptr, self._m.ptr, self._v.ptr, lo, hi, i1 - i0,
                             seg.ptr + i0 * 8, self._t_dev.ptr, float(self.lr), float(self.b1), float(self.b2),
                             float(self.eps), 0.0, i0, P if last else 0)
        if last:
            self.t += P

    def _mc_exchange_range(self, a, region, i0, i1, lo, hi, rank, world, last):
        """Parameters i0 .. i1-1 = arena elements [lo, hi): gradient reduce-scatter through the NVLink switch, Adam on
        this rank's 1/world share, all-gather of the new parameters -- one kernel (lg_mc_exchange_step).  Step
        counter and bias corrections as in ``_fused_step_range``."""
        P = len(self.parameters)
        if self._m is None:
            self._init_state(a)
        seg = a.segments(self.parameters)
        goff, poff, foff = region
        a.rt.api.mc_exchange_step(self._belief, goff, poff, foff, lo, hi, rank, world, self._m.ptr, self._v.ptr,
                                  i1 - i0, seg.ptr + i0 * 8, self._t_dev.ptr, float(self.lr), float(self.b1),
                                  float(self.b2), float(self.eps), 0.0, i0, P if last else 0)
        if last:
            self.t += P


class AdaBelief(Adam):
    """Adapting Stepsizes by the Belief in Observed Gradients (https://arxiv.org/abs/2010.07468)."""
    _belief = 1

    def _second_moment_input(self, grad, i):
        return grad - self.m[i]
