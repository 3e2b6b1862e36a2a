"""lightgrad_b200 -- a B200-native (sm_100a) CUDA tensor backend behind lightgrad's autograd API.

Same top-level surface as the reference package (lightgrad/__init__.py:1-6):
``autograd, loss, nn, optim`` sub-modules, ``Tensor``/``Gradients``/``no_grad``
and the initialiser shortcuts -- with ``Tensor = CudaTensor``.
There is no CPU fallback: creating a tensor without the compiled
``liblightgrad_b200.so`` or without a GPU raises.
"""
from . import autograd, loss, nn, optim
from .autograd import Tensor, CudaTensor, Gradients, no_grad

empty, zeros, ones = Tensor.empty, Tensor.zeros, Tensor.ones
uniform, xavier = Tensor.uniform, Tensor.xavier
from_numpy = Tensor.from_numpy
