from .tensor import CudaTensor  # noqa: F401
