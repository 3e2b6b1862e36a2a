"""ctypes binding of liblightgrad_b200.so (the C-ABI declared in include/lightgrad_b200.h).

The library is loaded lazily, on the first device operation.  There is no fallback: if the shared
object has not been built, or no sm_100 GPU is visible, the first tensor operation raises.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('LG_LIB') or os.path.normpath(os.path.join(_HERE, '..', '..', 'lib', 'liblightgrad_b200.so'))

# ---- enums mirrored from include/lightgrad_b200.h (tests/test_abi.py checks they agree) ----------
F32, F64, I32, I64, I16, U8, I8, BF16 = range(8)
DTYPE_CODE = {np.dtype(np.float32): F32, np.dtype(np.float64): F64, np.dtype(np.int32): I32,
              np.dtype(np.int64): I64, np.dtype(np.int16): I16, np.dtype(np.uint8): U8,
              np.dtype(np.int8): I8, np.dtype(np.bool_): U8}

EW = dict(COPY=0, NEG=1, SIN=2, COS=3, EXP=4, LOG=5, SIGMOID=6, TANH=7, RELU=8, GELU=9,
          ADD_S=10, MUL_S=11, RSUB_S=12, RDIV_S=13, POW_S=14, RPOW_S=15, SQRT=16, DIV_S=17, FILL=18,
          ADD=32, SUB=33, MUL=34, DIV=35, POW=36,
          SIN_BWD=40, COS_BWD=41, LOG_BWD=42, SIGMOID_BWD=43, TANH_BWD=44, RELU_BWD=45, GELU_BWD=46,
          POW_S_BWD=47, RPOW_S_BWD=48, RDIV_S_BWD=49, AXPY=50,
          DIV_BWD_B=64, POW_BWD_A=65, POW_BWD_B=66, EQ_MASK_MUL=67)
RED = dict(SUM=0, MAX=1, MIN=2)
GEMM_FP32_SIMT, GEMM_TF32_TC, GEMM_BF16_TC = 0, 1, 2
MAX_DIMS = 8

_i64p = C.POINTER(C.c_int64)
_vp = C.c_void_p


class GemmDesc(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        'M', 'N', 'K', 'batch0', 'batch1',
        'sa_b0', 'sa_b1', 'sa_m', 'sa_k',
        'sb_b0', 'sb_b1', 'sb_k', 'sb_n',
        'sc_b0', 'sc_b1', 'sc_m', 'sc_n')]


_SIGNATURES = {
    'lg_device_count': [C.POINTER(C.c_int)],
    'lg_init': [C.c_int],
    'lg_device': [C.POINTER(C.c_int)],
    'lg_device_props': [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)],
    'lg_sync': [],
    'lg_alloc': [C.c_size_t, C.POINTER(_vp)],
    'lg_free': [_vp],
    'lg_empty_cache': [],
    'lg_mem_stats': [C.POINTER(C.c_size_t)] * 3,
    'lg_memcpy_h2d': [_vp, _vp, C.c_size_t],
    'lg_memcpy_d2h': [_vp, _vp, C.c_size_t],
    'lg_memcpy_d2d': [_vp, _vp, C.c_size_t],
    'lg_memset': [_vp, C.c_int, C.c_size_t],
    'lg_host_alloc': [C.c_size_t, C.POINTER(_vp)],
    'lg_host_free': [_vp],
    'lg_event_create': [C.POINTER(_vp)],
    'lg_event_record': [_vp],
    'lg_event_sync': [_vp],
    'lg_event_elapsed_ms': [_vp, _vp, C.POINTER(C.c_float)],
    'lg_event_destroy': [_vp],
    'lg_launch_count': [C.POINTER(C.c_uint64)],
    'lg_profiler_range': [C.c_int],
    'lg_stream_delay_us': [C.c_uint64],
    'lg_graph_begin': [C.POINTER(C.c_int)],
    'lg_graph_end': [C.POINTER(_vp), C.POINTER(C.c_uint64)],
    'lg_graph_abort': [],
    'lg_comm_compute_begin': [],
    'lg_comm_compute_end': [],
    'lg_side_begin': [],
    'lg_side_end': [],
    'lg_side_join': [],
    'lg_graph_launch': [_vp, C.c_uint64],
    'lg_graph_destroy': [_vp],
    'lg_ew_flat': [C.c_int, C.c_int, _vp, _vp, _vp, _vp, C.c_int64, C.c_double],
    'lg_ew': [C.c_int, C.c_int, C.c_int, _i64p, _vp, _i64p, _vp, _i64p, _vp, _i64p, _vp, _i64p, C.c_double],
    'lg_ew_bwd2_flat': [C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int64],
    'lg_cast': [C.c_int, C.c_int, C.c_int, _i64p, _vp, _i64p, _vp, _i64p],
    'lg_reduce': [C.c_int, C.c_int, _vp, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_double],
    'lg_reduce_pitched': [C.c_int, C.c_int, _vp, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_int],
    'lg_gemm': [C.c_int, C.c_int, C.POINTER(GemmDesc), _vp, _vp, _vp, _vp, C.c_int],
    'lg_gemm_tc_supported': [C.c_int, C.c_int, C.POINTER(GemmDesc)],
    'lg_gemm_grouped': [C.c_int, C.c_int, C.POINTER(GemmDesc), C.c_int, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp),
                        C.POINTER(_vp), C.c_int],
    'lg_gemm_epilogue': [C.c_int, C.c_int, C.POINTER(GemmDesc), _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int64, C.c_double],
    'lg_gemm_sm_limit': [C.c_int],
    'lg_prof_gemm': [C.c_int],
    'lg_prof_gemm_read': [C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_double)],
    'lg_im2col': [C.c_int, C.c_int, C.c_int64, _i64p, _i64p, _i64p, _vp, _vp],
    'lg_col2im': [C.c_int, C.c_int, C.c_int64, _i64p, _i64p, _i64p, _vp, _vp],
    'lg_gather_rows': [C.c_int, C.c_int, _vp, C.c_int64, C.c_int64, _vp, C.c_int64, C.c_int64, _vp],
    'lg_scatter_add_rows': [C.c_int, C.c_int, _vp, C.c_int64, C.c_int64, _vp, C.c_int64, C.c_int64, _vp],
    'lg_scatter_set_rows': [C.c_int, C.c_int, _vp, C.c_int64, C.c_int64, _vp, C.c_int64, C.c_int64, _vp, C.c_double],
    'lg_index_linearize': [C.c_int, C.POINTER(_vp), C.POINTER(C.c_int), _i64p, _i64p, C.c_int64, _vp],
    'lg_softmax_fwd': [C.c_int, _vp, _vp, C.c_int64, C.c_int64, C.c_double],
    'lg_softmax_bwd': [C.c_int, _vp, _vp, _vp, C.c_int64, C.c_int64, C.c_double],
    'lg_cross_entropy_fwd': [C.c_int, C.c_int, _vp, C.c_int64, _vp, _vp, _vp, C.c_int64, C.c_int64],
    'lg_cross_entropy_bwd': [C.c_int, C.c_int, _vp, C.c_int64, _vp, _vp, _vp, _vp, C.c_int64, C.c_int64, C.c_int64],
    'lg_layernorm_fwd': [C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int64, C.c_double],
    'lg_add_layernorm_fwd': [C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int64, C.c_double],
    'lg_layernorm_bwd': [C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int64, C.c_int, _vp],
    'lg_attention_supported': [C.c_int, C.c_int64, C.c_int64],
    'lg_attention_fwd': [C.c_int, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_double, _vp, _vp],
    'lg_attention_bwd': [C.c_int, _vp, _vp, _vp, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_double, _vp,
                         _vp, _vp, _vp],
    'lg_sgd_step': [_vp, _vp, _vp, C.c_int64, C.c_double, C.c_double],
    'lg_adam_step': [C.c_int, _vp, _vp, _vp, _vp, C.c_int64, C.c_int, _vp, _vp,
                     C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64, C.c_int, C.c_int],
    'lg_nccl_unique_id': [_vp],
    'lg_nccl_init': [_vp, C.c_int, C.c_int],
    'lg_nccl_allreduce_f32': [_vp, C.c_int64, C.c_int, C.c_int],
    'lg_nccl_broadcast': [_vp, C.c_int64, C.c_int],
    'lg_nccl_wait': [],
    'lg_nccl_fork': [],
    'lg_nccl_destroy': [],
    'lg_mc_supported': [C.POINTER(C.c_int)],
    'lg_mc_region_bytes': [C.c_size_t, C.c_int] + [C.POINTER(C.c_size_t)] * 4,
    'lg_mc_create': [C.c_size_t, C.c_int, C.POINTER(C.c_int)],
    'lg_mc_import': [C.c_int, C.c_size_t, C.c_int],
    'lg_mc_add_device': [],
    'lg_mc_bind': [C.POINTER(_vp), C.POINTER(_vp)],
    'lg_mc_exchange_step': [C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int64, C.c_int64, C.c_int, C.c_int, _vp, _vp,
                            C.c_int, _vp, _vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                            C.c_int],
    'lg_mc_release': [],
    'lg_bucket_step': [C.c_int, _vp, _vp, _vp, _vp, C.c_int64, C.c_int64, C.c_int, _vp, _vp, C.c_double, C.c_double,
                       C.c_double, C.c_double, C.c_double, C.c_int, C.c_int],
    'lg_mc_trace_mark': [],
    'lg_mc_trace_read': [_vp, C.c_int, C.POINTER(C.c_int), C.c_int],
}

_lib = None


class _Checked(object):
    """Callable wrapper: raises RuntimeError with lg_last_error() on a non-zero status."""
    __slots__ = ('fn', 'name')

    def __init__(self, fn, name):
        self.fn, self.name = fn, name

    def __call__(self, *args):
        if self.fn(*args):
            msg = _lib.lg_last_error().decode()
            if msg.startswith('IndexError:'):
                # an out-of-range index / label met by a kernel, reported by the synchronisation that followed it
                raise IndexError(msg[len('IndexError:'):].strip())
            raise RuntimeError("%s: %s" % (self.name, msg))


class _Api(object):
    pass


api = None


def load():
    """Load the shared object and bind every entry point (idempotent)."""
    global _lib, api
    if api is not None:
        return api
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "lightgrad_b200: %s is missing -- build it with `python -m lightgrad_b200.build` "
            "(there is no CPU or PyTorch fallback)" % LIB_PATH)
    _lib = C.CDLL(LIB_PATH)
    _lib.lg_last_error.restype = C.c_char_p
    _lib.lg_last_error.argtypes = []
    _lib.lg_stream_handle.restype = _vp
    _lib.lg_stream_handle.argtypes = []
    a = _Api()
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(_lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
        if name in ('lg_gemm_tc_supported', 'lg_attention_supported'):
            setattr(a, name[3:], fn)       # predicates: the return value is the answer, not a status
        else:
            setattr(a, name[3:], _Checked(fn, name))
    a.raw = _lib
    api = a
    return a


_ready = False


def ensure_device(device=-1):
    """Bind this process to a GPU (LOCAL_RANK-th device by default).  Raises without one."""
    global _ready
    a = load()
    if not _ready:
        a.init(int(device))
        _ready = True
        from ..utils.profiler import Profiler
        Profiler.device_sync = synchronize
    return a


def synchronize():
    ensure_device().sync()
    _side_held.clear()
    _side_dirty.clear()


# ---- side stream (lg_side_begin / end / join) ----------------------------------------------------------
_side_held = []          # tensors the side stream's launches touch: kept alive until the join
_side_dirty = set()      # device addresses of gradients the side stream accumulates into


class side_stream(object):
    """``with side_stream(x, g): ...`` issues the block's launches on the side stream (they run concurrently
    with whatever the compute stream does next) and keeps the given tensors alive until ``side_join``."""

    def __init__(self, *hold, writes=()):
        self.hold, self.writes = hold, writes

    def __enter__(self):
        if side_prepare is not None:
            side_prepare(self.hold)      # anything the operands need staged happens on the compute stream first
        api.side_begin()
        _side_held.extend(self.hold)
        _side_held.extend(self.writes)
        for t in self.writes:
            _side_dirty.add(t.ptr)

    def __exit__(self, *exc):
        api.side_end()


side_prepare = None      # hook of the operator layer: called with the tensors a side-stream block will read


def side_join():
    """Compute stream waits for the side stream (no-op when nothing is outstanding)."""
    if api is not None and _ready:
        api.side_join()      # cheap when nothing is outstanding (the library tracks it)
    _side_held.clear()
    _side_dirty.clear()


def side_join_if_written(t):
    """Called before the compute stream updates ``t`` in place: orders it after side-stream writes to ``t``."""
    if _side_dirty and t.ptr in _side_dirty:
        side_join()


def launch_count():
    n = C.c_uint64(0)
    ensure_device().launch_count(C.byref(n))
    return n.value


def gemm_profile(mode):
    """0 / False: off; 1 / True: event pair around every lg_gemm launch; 2: count only, launch nothing."""
    ensure_device().prof_gemm(int(mode))


def gemm_profile_read():
    """(summed kernel ms, launches, algorithmic flops) of the lg_gemm launches since the last read."""
    ms, n, fl = C.c_double(0), C.c_uint64(0), C.c_double(0)
    ensure_device().prof_gemm_read(C.byref(ms), C.byref(n), C.byref(fl))
    return ms.value, n.value, fl.value


def mem_stats():
    a, b, c = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    ensure_device().mem_stats(C.byref(a), C.byref(b), C.byref(c))
    return {'in_use': a.value, 'reserved': b.value, 'peak_in_use': c.value}


def device_props():
    sm, ma, mi, mem = C.c_int(0), C.c_int(0), C.c_int(0), C.c_size_t(0)
    ensure_device().device_props(C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem))
    return {'sm_count': sm.value, 'cc': (ma.value, mi.value), 'total_mem': mem.value}


class Buffer(object):
    """Ref-counted device block from the library's caching allocator (analogue of PooledBuffer)."""
    __slots__ = ('ptr', 'nbytes', '_free', '_bf16', '_colsum', '__weakref__')

    def __init__(self, nbytes):
        self.ptr = 0
        self._bf16 = None          # bf16 staging copy of the whole block (bf16 tensor-core matmul mode), or None
        self._colsum = None        # (tensor, ptr, rows, cols, ld): column sums a producer kernel left for a matrix in here
        a = ensure_device()
        p = _vp()
        a.alloc(int(nbytes), C.byref(p))
        self.ptr = p.value or 0
        self.nbytes = int(nbytes)
        self._free = a.raw.lg_free

    def __del__(self):
        p = self.ptr
        if p:
            self.ptr = 0
            try:
                self._free(p)
            except Exception:
                pass


class ExternalBuffer(object):
    """Device memory this module does not own (a window of the NVLink multicast region): same face as Buffer, never
    freed here; ``keep`` holds whatever must outlive it."""
    __slots__ = ('ptr', 'nbytes', 'keep', '_bf16', '_colsum', '__weakref__')

    def __init__(self, ptr, nbytes, keep=None):
        self.ptr, self.nbytes, self.keep, self._bf16, self._colsum = int(ptr), int(nbytes), keep, None, None


class ArenaSlice(object):
    """A window into a parent Buffer (keeps the parent alive); used for flat parameter / gradient arenas."""
    __slots__ = ('ptr', 'nbytes', 'parent')

    def __init__(self, parent, byte_offset, nbytes):
        self.parent = parent
        self.ptr = parent.ptr + int(byte_offset)
        self.nbytes = int(nbytes)


class Event(object):
    __slots__ = ('h',)

    def __init__(self):
        h = _vp()
        ensure_device().event_create(C.byref(h))
        self.h = h.value

    def record(self):
        api.event_record(self.h)
        return self

    def synchronize(self):
        api.event_sync(self.h)

    def elapsed_ms(self, later):
        ms = C.c_float(0)
        api.event_elapsed_ms(self.h, later.h, C.byref(ms))
        return ms.value

    def __del__(self):
        try:
            if self.h and api is not None:
                api.raw.lg_event_destroy(self.h)
        except Exception:
            pass


class PinnedArray(object):
    """numpy view over page-locked host memory (H2D copies from it are truly asynchronous)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape)
        n = int(np.prod(self.shape)) * self.dtype.itemsize
        p = _vp()
        ensure_device().host_alloc(max(n, 1), C.byref(p))
        self.ptr = p.value
        buf = (C.c_char * max(n, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def __del__(self):
        try:
            if self.ptr and api is not None:
                api.raw.lg_host_free(self.ptr)
                self.ptr = 0
        except Exception:
            pass
