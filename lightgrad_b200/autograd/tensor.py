"""Backend-agnostic tensor base class.

Keeps the plugin contract of the reference (lightgrad/autograd/tensor.py:5-161):
a backend subclasses ``AbstractTensor``, implements ``dtype``, ``shape`` and the
initialisers, and attaches operators with ``register_op``.  A subclass defined
in module ``<pkg>.<backend>.tensor`` automatically gives every tensor a
``.<backend>()`` converter that round-trips through numpy (tensor.py:11-14,
154-161) -- ``lightgrad_b200.autograd.cuda.tensor.CudaTensor`` therefore adds
``.cuda()``.
"""
import numpy as np
from .grads import Gradients


class _TensorMeta(type):

    def __new__(mcls, name, bases, attrs):
        T = type.__new__(mcls, name, bases, attrs)
        mod = attrs.get('__module__')
        # skip the abstract base itself and classes synthesised at run time
        if mod is not None and mod != __name__:
            parts = mod.split('.')
            if len(parts) >= 2:
                AbstractTensor.register_backend(parts[-2], T)
        return T


class AbstractTensor(metaclass=_TensorMeta):

    def __init__(self, data, requires_grad=True):
        self._data = data
        self._grad = None
        self._requires_grad = requires_grad
        self._ctx = None
        # True while this object is the sole owner of a freshly computed buffer
        # (lets add_grad adopt it instead of copying)
        self._temp = False

    # -- graph bookkeeping ---------------------------------------------------
    def _set_ctx(self, ctx):
        assert ctx is None or isinstance(ctx, Function)
        self._ctx = ctx
        self._temp = False
        return self

    def _set_data(self, data):
        self._data = data
        return self

    def _mark_shared(self):
        self._temp = False

    def _drop_grad(self):
        self._grad = None

    def detach(self):
        # as in the reference: cuts the history in place and returns self
        self._ctx = None
        return self

    @property
    def ctx(self):
        return self._ctx

    @property
    def data(self):
        return self._data

    @property
    def grad(self):
        return self._grad

    @property
    def requires_grad(self):
        return self._requires_grad

    @property
    def dtype(self):
        raise NotImplementedError()

    @property
    def shape(self):
        raise NotImplementedError()

    def item(self):
        return self.numpy().item()

    def numel(self):
        n = 1
        for s in self.shape:
            n *= s
        return int(n)

    # -- initialisers (backend supplies these) -------------------------------
    @staticmethod
    def empty(shape, requires_grad=True):
        raise NotImplementedError()

    @staticmethod
    def zeros(shape, requires_grad=True):
        raise NotImplementedError()

    @staticmethod
    def ones(shape, requires_grad=True):
        raise NotImplementedError()

    @staticmethod
    def uniform(low, high, shape, requires_grad=True):
        raise NotImplementedError()

    @staticmethod
    def from_numpy(a, requires_grad=True):
        raise NotImplementedError()

    @classmethod
    def xavier(cls, shape, requires_grad=True):
        # uniform(-1, 1) / sqrt(numel), scaled in place (tensor.py:85-89)
        t = cls.uniform(-1, 1, shape=shape, requires_grad=requires_grad)
        t /= float(np.sqrt(t.numel()))
        return t.detach()

    def copy(self, requires_grad=True):
        raise NotImplementedError()

    def numpy(self):
        raise NotImplementedError()

    # -- gradients -----------------------------------------------------------
    def backward(self, allow_fill=False):
        if self._ctx is None:
            return
        shp = self.shape
        if shp == (1,) or len(shp) == 0 or allow_fill:
            self._grad = self.__class__.ones(shp, requires_grad=False)
        else:
            raise RuntimeError("Can only backpropagate from item tensors!")
        Gradients.backward(self._ctx, self._grad)

    def add_grad(self, grad):
        if not self._requires_grad:
            return
        Gradients._depth += 1
        try:
            if self._grad is None:
                if grad._temp:
                    # freshly computed and unshared: adopt instead of copying
                    grad._temp = False
                    grad._requires_grad = False
                    self._grad = grad
                else:
                    self._grad = grad.copy(requires_grad=False)
                    self._grad._temp = False
            else:
                self._grad += grad
        finally:
            d = Gradients._depth - 1
            Gradients._depth = d if d > 0 else 0

    def zero_grad(self, traverse_graph=False):
        if self._requires_grad:
            if self._grad is None:
                self._grad = self.__class__.zeros(self.shape, requires_grad=False)
            else:
                self._grad.fill(0)
        if traverse_graph and self._ctx is not None:
            seen = {id(self)}
            todo = [self._ctx]
            while todo:
                c = todo.pop()
                for t in c.parent_tensors:
                    if id(t) in seen:
                        continue
                    seen.add(id(t))
                    if t._requires_grad:
                        if t._grad is None:
                            t._grad = t.__class__.zeros(t.shape, requires_grad=False)
                        else:
                            t._grad.fill(0)
                    if t._ctx is not None:
                        todo.append(t._ctx)

    # -- registries ----------------------------------------------------------
    @classmethod
    def register_op(cls, name=None, op=None, overwrite=False):
        if op is None:
            # decorator form: @T.register_op(), @T.register_op("name")
            return lambda o: cls.register_op(name if name is not None else o.__name__, o, overwrite=overwrite)
        if not (isinstance(op, type) and issubclass(op, Function)):
            raise TypeError("Operators must inherit from Function! (%s)" % getattr(op, '__name__', op))
        if not overwrite and hasattr(cls, name):
            raise RuntimeError("Function %s already registered to %s!" % (name, cls.__name__))

        def dispatch(self, *args, **kwargs):
            return op(self, *args, **kwargs)
        dispatch.__name__ = name
        dispatch.__doc__ = op.__doc__
        dispatch.op = op
        setattr(cls, name, dispatch)
        return op

    @staticmethod
    def register_backend(name, Tensor_cls):
        if not issubclass(Tensor_cls, AbstractTensor):
            raise TypeError("Backend tensors must inherit from Tensor! (%s)" % Tensor_cls.__name__)

        def convert(t, *args, **kwargs):
            return Tensor_cls.from_numpy(t.numpy(), *args, **kwargs)
        convert.__name__ = name
        setattr(AbstractTensor, name, convert)


from .func import Function  # noqa: E402
from . import ops  # noqa: E402,F401  (registers the generic composite operators)
