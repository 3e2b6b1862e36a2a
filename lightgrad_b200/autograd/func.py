"""Operator protocol: ``Function`` and ``WrapperFunction``.

Same contract as the reference (lightgrad/autograd/func.py:5-106):

* ``Op(*args, **kwargs)`` does not build an ``Op`` instance for the caller; it
  builds a context object, runs ``forward`` with graph recording suspended
  and returns the output tensor, whose ``ctx`` becomes the context when
  recording is on (func.py:11-29).
* positional arguments are the graph parents; keyword arguments must not
  carry tensors that require gradients (func.py:13).
* all tensor operands must share one tensor class (func.py:18-20).
* ``backward`` returns one gradient per positional parent; gradients that are
  broadcast-shaped are summed back to the parent's shape (func.py:50-56).
* ``WrapperFunction`` differentiates a composite by walking the graph that
  its ``forward`` recorded, with the operands masked as leaves
  (func.py:71-106).
"""
from .grads import Gradients
from .utils.profiler import Tracker, Profiler


class _FunctionMeta(type):

    def _run(cls, f, args, kwargs):
        # primitive ops: forward is opaque to the graph
        Gradients._depth += 1
        try:
            return f.forward(*args, **kwargs)
        finally:
            d = Gradients._depth - 1
            Gradients._depth = d if d > 0 else 0

    def __call__(cls, *args, **kwargs):
        first = None
        for a in args:
            if isinstance(a, AbstractTensor):
                if first is None:
                    first = a.__class__
                elif not isinstance(a, first):
                    raise AssertionError("All Tensors must be of the same type! %s" % str(
                        tuple(t.__class__.__name__ for t in args if isinstance(t, AbstractTensor))))
        for v in kwargs.values():
            if isinstance(v, AbstractTensor):
                assert not v.requires_grad, "tensors passed by keyword must not require gradients"
                if first is None:
                    first = v.__class__
                else:
                    assert isinstance(v, first), "All Tensors must be of the same type!"
        f = object.__new__(cls)
        f.__init__(*args)
        if Profiler._active_profilers:
            with Tracker(cls.__name__):
                out = cls._run(f, args, kwargs)
        else:
            out = cls._run(f, args, kwargs)
        assert isinstance(out, AbstractTensor)
        if Gradients._depth == 0:
            out._set_ctx(f)
        return out


class Function(object, metaclass=_FunctionMeta):
    # a backward may add a parent's gradient straight into ``parent.grad`` (e.g. a GEMM epilogue that
    # accumulates into the optimizer's gradient arena) and return this marker in that slot instead
    ACCUMULATED = type('Accumulated', (), {'__repr__': lambda self: 'Function.ACCUMULATED'})()

    def __init__(self, *parents):
        self._parents = parents
        self._saved = ()

    @property
    def parent_tensors(self):
        """Parents that take part in differentiation."""
        return (t for t in self._parents if isinstance(t, AbstractTensor) and t.requires_grad)

    def _deliver(self, in_grads):
        in_grads = in_grads if isinstance(in_grads, tuple) else (in_grads,)
        for t, g in zip(self._parents, in_grads):
            if not (isinstance(t, AbstractTensor) and t.requires_grad):
                continue
            if g is Function.ACCUMULATED:
                continue
            assert g is not None
            gs, ts = g.shape, t.shape
            if gs != ts:
                # undo broadcasting: sum over the axes numpy expanded
                assert len(gs) >= len(ts), "Cannot unbroadcast shapes %s and %s" % (ts, gs)
                lead = len(gs) - len(ts)
                axes = tuple(range(lead)) + tuple(
                    lead + i for i, (x, y) in enumerate(zip(ts, gs[lead:])) if x != y)
                g = g.sum(axis=axes, keepdims=True)
                g = g.reshape(*g.shape[lead:])
            assert g.shape == ts
            t.add_grad(g)

    def _backpropagate(self, out_grad):
        if Profiler._active_profilers:
            with Tracker(self.__class__.__name__, backward=True):
                self._deliver(self.backward(out_grad))
        else:
            self._deliver(self.backward(out_grad))

    def save_for_backward(ctx, *args):
        for a in args:
            if isinstance(a, AbstractTensor):
                a._mark_shared()
        ctx._saved += tuple(args)

    def get_saved_tensors(self):
        return self._saved

    def forward(ctx, t, *args, **kwargs):
        raise NotImplementedError()

    def backward(ctx, out_grad):
        raise RuntimeError("Cannot Backward through %s!" % ctx.__class__.__name__)


class _WrapperMeta(_FunctionMeta):

    def _run(cls, f, args, kwargs):
        # composite ops: record the inner graph, then hide it behind ``f``
        out = f.forward(*args, **kwargs)
        f._set_internal_ctx(out.ctx)
        return out


class WrapperFunction(Function, metaclass=_WrapperMeta):
    """Function whose gradient comes from back-propagating through the graph its forward built."""

    def __init__(self, *parents):
        Function.__init__(self, *parents)
        self._inner = None

    def _set_internal_ctx(self, ctx):
        self._inner = ctx

    def _walk_inner(self, out_grad):
        if self._inner is None:
            return
        borders = [(p, p.ctx) for p in self.parent_tensors]
        for p, _ in borders:
            p._set_ctx(None)
        try:
            Gradients.backward(self._inner, out_grad, retain=False)
        finally:
            for p, c in borders:
                p._set_ctx(c)

    def _backpropagate(self, out_grad):
        if Profiler._active_profilers:
            with Tracker(self.__class__.__name__, backward=True):
                self._walk_inner(out_grad)
        else:
            self._walk_inner(out_grad)

    def forward(ctx, *args, **kwargs):
        raise NotImplementedError()

    @staticmethod
    def from_function(fn):
        """Decorator: turn ``fn(*tensors, **kw)`` into a WrapperFunction class of the same name."""
        return type(fn.__name__, (WrapperFunction,), {
            'forward': (lambda ctx, *args, **kwargs: fn(*args, **kwargs)),
            '__doc__': fn.__doc__,
        })


from .tensor import AbstractTensor  # noqa: E402  (circular by design, as in the reference)
