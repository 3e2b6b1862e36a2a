"""The C-ABI boundary: the shared library loads without a GPU, exports every symbol that
include/lightgrad_b200.h declares, the python enum mirrors agree with the header, and the product
fails loudly (no CPU fallback) when no device is present."""
import ctypes
import os
import re
import numpy as np
import pytest
from lightgrad_b200.autograd.cuda import runtime as rt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, 'include', 'lightgrad_b200.h')).read()


def _declared_functions():
    body = re.sub(r'/\*.*?\*/', '', HEADER, flags=re.S)
    return sorted(set(re.findall(r'\b(lg_[a-z0-9_]+)\s*\(', body)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(rt.LIB_PATH), "build the library first: python -m lightgrad_b200.build"
    lib = ctypes.CDLL(rt.LIB_PATH)
    names = _declared_functions()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    declared = set(_declared_functions())
    bound = set(rt._SIGNATURES) | {'lg_last_error', 'lg_stream_handle'}
    assert declared == bound, (sorted(declared - bound), sorted(bound - declared))


def test_enum_mirrors_match_header():
    enums = dict((k, int(v)) for k, v in re.findall(r'\b(LG_[A-Z0-9_]+)\s*=\s*(\d+)', HEADER))
    for name, code in rt.EW.items():
        assert enums['LG_EW_' + name] == code, name
    for name, code in rt.RED.items():
        assert enums['LG_RED_' + name] == code
    for name, code in (('F32', rt.F32), ('F64', rt.F64), ('I32', rt.I32), ('I64', rt.I64), ('I16', rt.I16),
                       ('U8', rt.U8), ('I8', rt.I8), ('BF16', rt.BF16)):
        assert enums['LG_' + name] == code
    assert (enums['LG_GEMM_FP32_SIMT'], enums['LG_GEMM_TF32_TC'], enums['LG_GEMM_BF16_TC']) == \
        (rt.GEMM_FP32_SIMT, rt.GEMM_TF32_TC, rt.GEMM_BF16_TC)
    assert int(re.search(r'#define LG_MAX_DIMS (\d+)', HEADER).group(1)) == rt.MAX_DIMS
    assert ctypes.sizeof(rt.GemmDesc) == 17 * 8


def test_no_cpu_fallback_without_a_device():
    lib = ctypes.CDLL(rt.LIB_PATH)
    n = ctypes.c_int(0)
    lib.lg_device_count(ctypes.byref(n))
    if n.value > 0:
        pytest.skip("a GPU is visible here")
    lib.lg_last_error.restype = ctypes.c_char_p
    assert lib.lg_init(-1) != 0
    assert b"no CPU fallback" in lib.lg_last_error()
    import lightgrad_b200 as light
    with pytest.raises(RuntimeError):
        light.CudaTensor.from_numpy(np.ones(3, dtype=np.float32))


@pytest.mark.gpu
def test_runtime_roundtrip_and_allocator(cuda):
    from lightgrad_b200 import CudaTensor
    x = np.random.RandomState(0).randn(1000, 37).astype(np.float32)
    t = CudaTensor.from_numpy(x)
    np.testing.assert_array_equal(t.numpy(), x)
    before = cuda.mem_stats()
    for _ in range(10):
        CudaTensor.empty((1 << 20,))
    after = cuda.mem_stats()
    assert after['reserved'] - before['reserved'] <= 8 << 20, "freed blocks must be reused"
    props = cuda.device_props()
    assert props['cc'][0] == 10 and props['sm_count'] >= 100


def test_test_double_covers_every_entry_point():
    # the numpy stand-in of the C-ABI (tests/fake_device.py) must follow the header: a new entry point without a
    # stand-in would silently drop its host logic from the CPU suite
    from lightgrad_b200.autograd.cuda import runtime as rt
    from tests import fake_device as fd
    missing = [n for n in rt._SIGNATURES if not hasattr(fd.FakeDevice, n[3:])]
    assert not missing, missing
