from .tensor import CpuTensor  # noqa: F401
