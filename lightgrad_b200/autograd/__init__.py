"""Autograd core + the sm_100a CUDA tensor backend (layout of lightgrad/autograd/__init__.py)."""
from .grads import Gradients
from .func import Function, WrapperFunction
from .tensor import AbstractTensor
from .cuda import CudaTensor

# the only backend this package ships is the B200 one
Tensor = CudaTensor
no_grad = Gradients.no_grad
