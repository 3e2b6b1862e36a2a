"""TEST DOUBLE of the C-ABI (include/lightgrad_b200.h) -- host-logic tests only.

The python side of the cuda backend (views, strides, broadcasting, index plans, reshape rules,
optimizer arenas, data-parallel bucketing) is several hundred lines of bookkeeping that can be
wrong without any kernel being wrong.  This module stands in for liblightgrad_b200.so with numpy
acting on host memory, so that bookkeeping can be exercised in the CPU-only container
(`pytest -m "not gpu"`).  It lives under tests/, is installed only by the `fake_device` fixture, and
is never imported by the product package: without the real library and a GPU the product raises.
It makes no parity claim -- the parity tests proper are the `-m gpu` tests.
"""
import ctypes as C
import numpy as np
from lightgrad_b200.autograd.cuda import runtime as rt

NP = {rt.F32: np.float32, rt.F64: np.float64, rt.I32: np.int32, rt.I64: np.int64, rt.I16: np.int16,
      rt.U8: np.uint8, rt.I8: np.int8}


def _lst(a, n=None):
    if a is None:
        return None
    return [int(a[i]) for i in range(len(a) if n is None else n)]


def _arr(ptr, dtype, shape, strides=None):
    """numpy view of host memory at ``ptr``; strides in elements."""
    dt = np.dtype(NP[dtype] if not isinstance(dtype, (type, np.dtype)) else dtype)
    shape = tuple(int(s) for s in shape)
    if strides is None:
        strides, acc = [], 1
        for s in reversed(shape):
            strides.append(acc)
            acc *= s
        strides = list(reversed(strides))
    base = np.frombuffer((C.c_char * dt.itemsize).from_address(int(ptr)), dtype=dt)
    return np.lib.stride_tricks.as_strided(base, shape=shape, strides=[int(s) * dt.itemsize for s in strides])


def _gelu(x):
    return (0.5 * x) * (1.0 + np.tanh((x * 0.7978845608) * (1.0 + (0.044715 * x) * x)))


def _gelu_bwd(x, g):
    c1, c2 = 0.7978845608, 0.044715
    u = (x * c1) * (1.0 + c2 * x * x)
    t = np.tanh(u)
    return (0.5 * (1.0 + t) + (0.5 * x) * (1.0 - t * t) * (c1 * (1.0 + 3.0 * c2 * x * x))) * g


E = rt.EW
EW1 = {E['COPY']: lambda a, s: a, E['NEG']: lambda a, s: -a, E['SIN']: lambda a, s: np.sin(a),
       E['COS']: lambda a, s: np.cos(a), E['EXP']: lambda a, s: np.exp(a), E['LOG']: lambda a, s: np.log(a),
       E['SIGMOID']: lambda a, s: 1 / (1 + np.exp(-a)), E['TANH']: lambda a, s: np.tanh(a),
       E['RELU']: lambda a, s: np.maximum(a, 0), E['GELU']: lambda a, s: _gelu(a),
       E['ADD_S']: lambda a, s: a + s, E['MUL_S']: lambda a, s: a * s, E['RSUB_S']: lambda a, s: s - a,
       E['RDIV_S']: lambda a, s: s / a, E['POW_S']: lambda a, s: a ** s, E['RPOW_S']: lambda a, s: s ** a,
       E['SQRT']: lambda a, s: np.sqrt(a), E['DIV_S']: lambda a, s: a / s,
       E['FILL']: lambda a, s: np.full_like(a, s)}
EW2 = {E['ADD']: lambda a, b, s: a + b, E['SUB']: lambda a, b, s: a - b, E['MUL']: lambda a, b, s: a * b,
       E['DIV']: lambda a, b, s: a / b, E['POW']: lambda a, b, s: a ** b,
       E['SIN_BWD']: lambda a, b, s: np.cos(a) * b, E['COS_BWD']: lambda a, b, s: -np.sin(a) * b,
       E['LOG_BWD']: lambda a, b, s: (1 / a) * b, E['SIGMOID_BWD']: lambda a, b, s: a * (1 - a) * b,
       E['TANH_BWD']: lambda a, b, s: (1 - a * a) * b, E['RELU_BWD']: lambda a, b, s: b * (a >= 0),
       E['GELU_BWD']: lambda a, b, s: _gelu_bwd(a, b),
       E['POW_S_BWD']: lambda a, b, s: s * a ** (s - 1) * b, E['RPOW_S_BWD']: lambda a, b, s: b * a * np.log(s),
       E['RDIV_S_BWD']: lambda a, b, s: -s / (a * a) * b, E['AXPY']: lambda a, b, s: a + s * b}
EW3 = {E['DIV_BWD_B']: lambda a, b, c, s: -a / (b * b) * c, E['POW_BWD_A']: lambda a, b, c, s: b * a ** (b - 1) * c,
       E['POW_BWD_B']: lambda a, b, c, s: c * b * np.log(a), E['EQ_MASK_MUL']: lambda a, b, c, s: c * (a == b)}


class _Raw(object):
    def __init__(self, dev):
        self.dev = dev

    def lg_free(self, p):
        self.dev.free(p)
        return 0

    def lg_event_destroy(self, h):
        return 0

    def lg_host_free(self, p):
        self.dev.free(p)
        return 0


class FakeDevice(object):
    """Implements every entry point that runtime._SIGNATURES names (minus the lg_ prefix)."""

    def __init__(self):
        self.blocks = {}
        self.launches = 0
        self.raw = _Raw(self)
        self.mode_log = []

    # ---- runtime
    def init(self, device): pass
    def device_count(self, ref): ref._obj.value = 1
    def device(self, ref): ref._obj.value = 0
    def host_free(self, p): pass
    def sync(self): pass
    def empty_cache(self): pass
    def profiler_range(self, start): pass
    def stream_delay_us(self, us): pass

    # graphs: the double executes eagerly while 'capturing' and re-runs the recorded python thunk on replay
    def graph_begin(self, pool): pool._obj.value = 1
    def graph_end(self, exec_ref, n_ref): exec_ref._obj.value, n_ref._obj.value = 1, 0
    def graph_abort(self): pass
    def graph_launch(self, h, n): raise RuntimeError('the fake device cannot replay graphs')
    def graph_destroy(self, h): pass

    def alloc(self, nbytes, ref):
        buf = np.zeros(int(nbytes) + 64, dtype=np.uint8)
        addr = (buf.ctypes.data + 63) // 64 * 64
        self.blocks[addr] = buf
        ref._obj.value = addr
    host_alloc = alloc

    def free(self, p):
        self.blocks.pop(int(p), None)

    def mem_stats(self, a, b, c):
        tot = sum(v.nbytes for v in self.blocks.values())
        a._obj.value = b._obj.value = c._obj.value = tot

    def device_props(self, sm, ma, mi, mem):
        sm._obj.value, ma._obj.value, mi._obj.value, mem._obj.value = 148, 10, 0, 180 << 30

    def launch_count(self, ref):
        ref._obj.value = self.launches

    def memcpy_h2d(self, dst, src, n):
        C.memmove(int(dst), int(src), int(n))
    memcpy_d2h = memcpy_d2d = memcpy_h2d

    def memset(self, dst, byte, n):
        C.memset(int(dst), int(byte), int(n))

    def event_create(self, ref): ref._obj.value = 1
    def event_destroy(self, h): pass
    def event_record(self, h): pass
    def event_sync(self, h): pass
    def event_elapsed_ms(self, a, b, ref): ref._obj.value = 0.0

    # ---- elementwise
    def ew_flat(self, op, dt, a, b, c, out, n, alpha):
        self.ew(op, dt, 1, [n], a, None, b, None, c, None, out, None, alpha)

    def ew(self, op, dt, nd, shape, a, sa, b, sb, c, sc, out, so, alpha):
        self.launches += 1
        shape = _lst(shape, nd)
        T = NP[dt]
        alpha = T(alpha)
        A = _arr(a, dt, shape, _lst(sa, nd))
        O = _arr(out, dt, shape, _lst(so, nd))
        with np.errstate(all='ignore'):
            if op in EW1:
                r = EW1[op](A, alpha)
            elif op in EW2:
                r = EW2[op](A, _arr(b, dt, shape, _lst(sb, nd)), alpha)
            else:
                r = EW3[op](A, _arr(b, dt, shape, _lst(sb, nd)), _arr(c, dt, shape, _lst(sc, nd)), alpha)
        O[...] = np.asarray(r, dtype=T)

    def ew_bwd2_flat(self, kind, dt, a, b, g, da, db, n):
        self.launches += 1
        A, B, G = (_arr(p, dt, [n]) for p in (a, b, g))
        if kind == 0:
            ra, rb = G * B, A * G
        else:
            ra, rb = G / B, -A / (B * B) * G
        _arr(da, dt, [n])[...] = ra
        _arr(db, dt, [n])[...] = rb

    def cast(self, sd, dd, nd, shape, src, ss, dst, ds):
        self.launches += 1
        shape = _lst(shape, nd)
        _arr(dst, dd, shape, _lst(ds, nd))[...] = _arr(src, sd, shape, _lst(ss, nd)).astype(NP[dd])

    # ---- reductions
    def reduce(self, op, dt, x, out, outer, red, inner, scale):
        self.launches += 1
        X = _arr(x, dt, [outer, red, inner])
        fn = {0: np.sum, 1: np.max, 2: np.min}[op]
        r = fn(X, axis=1)
        if op == 0:
            r = r * NP[dt](scale)
        _arr(out, dt, [outer, inner])[...] = r

    def reduce_pitched(self, op, dt, x, out, outer, red, inner, ld, scale, accumulate):
        self.launches += 1
        X = _arr(x, dt, [outer, red, inner], [red * ld, ld, 1])
        fn = {0: np.sum, 1: np.max, 2: np.min}[op]
        r = fn(X, axis=1)
        if op == 0:
            r = r * NP[dt](scale)
        O = _arr(out, dt, [outer, inner])
        O[...] = (O + r) if accumulate else r

    # ---- matmul
    def gemm(self, mode, dt, dref, a, b, c, bias, accumulate):
        self.launches += 1
        self.mode_log.append(mode)
        d = dref._obj
        A = _arr(a, dt, [d.batch0, d.batch1, d.M, d.K], [d.sa_b0, d.sa_b1, d.sa_m, d.sa_k])
        B = _arr(b, dt, [d.batch0, d.batch1, d.K, d.N], [d.sb_b0, d.sb_b1, d.sb_k, d.sb_n])
        Cm = _arr(c, dt, [d.batch0, d.batch1, d.M, d.N], [d.sc_b0, d.sc_b1, d.sc_m, d.sc_n])
        r = A @ B
        if bias:
            r = r + _arr(bias, dt, [d.N])
        if accumulate:
            r = r + Cm
        Cm[...] = r

    def gemm_grouped(self, mode, dt, dref, groups, a, b, c, bias, accumulate):
        seen = []
        for g in range(groups):
            again = c[g] in seen          # groups naming one result: C = sum_g A_g B_g (+ bias once)
            self.gemm(mode, dt, dref, a[g], b[g], c[g], bias[g] if bias and not again else None,
                      1 if (accumulate or again) else 0)
            seen.append(c[g])

    def gemm_epilogue(self, mode, dt, dref, a, b, c, bias, epi, aux, aux_ld, alpha):
        d = dref._obj
        self.gemm(mode, dt, dref, a, b, c, bias, 0)
        f = NP[dt]
        if epi >= 3:
            Cm = _arr(c, dt, [d.batch0, d.batch1, d.M, d.N], [d.sc_b0, d.sc_b1, d.sc_m, 1])
            if epi == 3:
                X = Cm * f(alpha)
                e = np.exp(X - X.max(axis=-1, keepdims=True))
                Cm[...] = e / e.sum(axis=-1, keepdims=True)
            else:
                P = _arr(aux, dt, [d.batch0, d.batch1, d.M, d.N], [d.sc_b0, d.sc_b1, aux_ld, 1])
                Cm[...] = P * (Cm - (P * Cm).sum(axis=-1, keepdims=True)) * f(alpha)
            return
        Cm = _arr(c, dt, [d.M, d.N], [d.sc_m, 1])
        X = _arr(aux, dt, [d.M, d.N], [aux_ld, 1])
        c1, c2 = f(0.7978845608), f(0.044715)
        if epi == 1:
            inner = (Cm * c1) * (f(1.0) + (c2 * Cm) * Cm)
            X[...] = (f(0.5) * Cm) * (f(1.0) + np.tanh(inner))
        else:
            x2 = X * X
            t = np.tanh((X * c1) * (f(1.0) + c2 * x2))
            du = c1 * (f(1.0) + f(3.0) * c2 * x2)
            Cm[...] = (f(0.5) * (f(1.0) + t) + (f(0.5) * X) * (f(1.0) - t * t) * du) * Cm

    def gemm_sm_limit(self, n): pass

    def comm_compute_begin(self): pass

    def comm_compute_end(self): pass

    def side_begin(self): pass

    def side_end(self): pass

    def side_join(self): pass

    def gemm_tc_supported(self, mode, dt, dref):
        return 1 if mode != rt.GEMM_FP32_SIMT and dt == rt.F32 else 0

    def prof_gemm(self, enable): pass

    def prof_gemm_read(self, ms, n, fl):
        ms._obj.value, n._obj.value, fl._obj.value = 0.0, 0, 0.0

    # ---- indexing
    def gather_rows(self, dt, idt, src, n_src, row_stride, idx, n_idx, row_len, out):
        self.launches += 1
        I = _arr(idx, idt, [n_idx]).astype(np.int64)
        I = np.where(I < 0, I + n_src, I)
        O = _arr(out, dt, [n_idx, row_len])
        for i, r in enumerate(I):
            O[i] = _arr(int(src) + int(r) * row_stride * np.dtype(NP[dt]).itemsize, dt, [row_len])

    def scatter_add_rows(self, dt, idt, dst, n_dst, row_stride, idx, n_idx, row_len, src):
        self.launches += 1
        I = _arr(idx, idt, [n_idx]).astype(np.int64)
        I = np.where(I < 0, I + n_dst, I)
        S = _arr(src, dt, [n_idx, row_len])
        for i, r in enumerate(I):
            _arr(int(dst) + int(r) * row_stride * np.dtype(NP[dt]).itemsize, dt, [row_len])[...] += S[i]

    def scatter_set_rows(self, dt, idt, dst, n_dst, row_stride, idx, n_idx, row_len, src, value):
        self.launches += 1
        I = _arr(idx, idt, [n_idx]).astype(np.int64)
        I = np.where(I < 0, I + n_dst, I)
        S = _arr(src, dt, [n_idx, row_len]) if src else None
        for i, r in enumerate(I):
            _arr(int(dst) + int(r) * row_stride * np.dtype(NP[dt]).itemsize, dt, [row_len])[...] = \
                S[i] if S is not None else value

    def index_linearize(self, k, ptrs, dts, sizes, strides, n, lin):
        self.launches += 1
        off = np.zeros(n, dtype=np.int64)
        for j in range(k):
            I = _arr(ptrs[j], dts[j], [n]).astype(np.int64)
            I = np.where(I < 0, I + sizes[j], I)
            off += I * strides[j]
        _arr(lin, rt.I64, [n])[...] = off

    # ---- fused
    def softmax_fwd(self, dt, x, y, rows, cols, scale):
        self.launches += 1
        X = _arr(x, dt, [rows, cols]) * NP[dt](scale)
        e = np.exp(X - X.max(axis=1, keepdims=True))
        _arr(y, dt, [rows, cols])[...] = e / e.sum(axis=1, keepdims=True)

    def softmax_bwd(self, dt, y, g, dx, rows, cols, scale):
        self.launches += 1
        Y, G = _arr(y, dt, [rows, cols]), _arr(g, dt, [rows, cols])
        _arr(dx, dt, [rows, cols])[...] = NP[dt](scale) * (Y * (G - (Y * G).sum(axis=1, keepdims=True)))

    def cross_entropy_fwd(self, dt, idt, x, ld, lab, loss_rows, lse, rows, cols):
        self.launches += 1
        X, L = _arr(x, dt, [rows, cols], [ld, 1]), _arr(lab, idt, [rows]).astype(np.int64)
        m = X.max(axis=1)
        l = m + np.log(np.exp(X - m[:, None]).sum(axis=1))
        _arr(lse, dt, [rows])[...] = l
        _arr(loss_rows, dt, [rows])[...] = l - X[np.arange(rows), L]

    def cross_entropy_bwd(self, dt, idt, x, ld, lab, lse, gs, dx, ld_dx, rows, cols):
        self.launches += 1
        X, L = _arr(x, dt, [rows, cols], [ld, 1]), _arr(lab, idt, [rows]).astype(np.int64)
        p = np.exp(X - _arr(lse, dt, [rows])[:, None])
        p[np.arange(rows), L] -= 1
        _arr(dx, dt, [rows, cols], [ld_dx, 1])[...] = p / NP[dt](rows) * _arr(gs, dt, [1])[0]

    def layernorm_fwd(self, dt, x, w, b, y, mean, rstd, rows, cols, eps):
        self.launches += 1
        X = _arr(x, dt, [rows, cols])
        mu = X.mean(axis=1, keepdims=True)
        var = ((X - mu) ** 2).mean(axis=1, keepdims=True)
        rs = 1 / np.sqrt(var + NP[dt](eps))
        _arr(y, dt, [rows, cols])[...] = (X - mu) * rs * _arr(w, dt, [cols]) + _arr(b, dt, [cols])
        _arr(mean, dt, [rows])[...] = mu[:, 0]
        _arr(rstd, dt, [rows])[...] = rs[:, 0]

    def add_layernorm_fwd(self, dt, a, b, s, w, bias, y, mean, rstd, rows, cols, eps):
        _arr(s, dt, [rows, cols])[...] = _arr(a, dt, [rows, cols]) + _arr(b, dt, [rows, cols])
        self.layernorm_fwd(dt, s, w, bias, y, mean, rstd, rows, cols, eps)

    def layernorm_bwd(self, dt, x, w, mean, rstd, g, dx, dw, db, rows, cols, accumulate):
        self.launches += 1
        X, G, W = _arr(x, dt, [rows, cols]), _arr(g, dt, [rows, cols]), _arr(w, dt, [cols])
        mu, rs = _arr(mean, dt, [rows])[:, None], _arr(rstd, dt, [rows])[:, None]
        xh = (X - mu) * rs
        dxh = G * W
        _arr(dx, dt, [rows, cols])[...] = rs * (dxh - dxh.mean(axis=1, keepdims=True)
                                                - xh * (dxh * xh).mean(axis=1, keepdims=True))
        DW, DB = _arr(dw, dt, [cols]), _arr(db, dt, [cols])
        DW[...] = (DW if accumulate else 0) + (G * xh).sum(axis=0)
        DB[...] = (DB if accumulate else 0) + G.sum(axis=0)

    # ---- optimizers
    def sgd_step(self, p, g, d, n, lr, mom):
        self.launches += 1
        P, G = _arr(p, rt.F32, [n]), _arr(g, rt.F32, [n])
        delta = np.float32(-lr) * G
        if d:
            D = _arr(d, rt.F32, [n])
            delta = delta + np.float32(mom) * D
            D[...] = delta
        P[...] += delta

    def adam_step(self, belief, p, g, m, v, n, n_seg, seg_end, t_dev, lr, b1, b2, eps, seg_base, seg_offset,
                  t_advance):
        self.launches += 1
        tcount = _arr(t_dev, rt.I64, [1])
        t0 = int(tcount[0])
        tcount[0] = t0 + t_advance
        P, G, M, V = (_arr(q, rt.F32, [n]) for q in (p, g, m, v))
        ends = _arr(seg_end, rt.I64, [n_seg])
        seg = np.searchsorted(ends, np.arange(n) + seg_base, side='right')
        t = (t0 + seg_offset + seg + 1).astype(np.float64)
        d1 = (1.0 - np.power(b1, t)).astype(np.float32)
        d2 = (1.0 - np.power(b2, t)).astype(np.float32)
        M[...] = np.float32(b1) * M + np.float32(1.0 - b1) * G
        r = (G - M) if belief else G
        V[...] = np.float32(b2) * V + np.float32(1.0 - b2) * (r * r)
        P[...] += np.float32(-lr) * (M / d1) / (np.sqrt(V / d2) + np.float32(eps))

    # ---- collectives: single-process stand-ins (identity)
    def nccl_unique_id(self, p): pass
    def nccl_init(self, p, world, rank): pass
    def nccl_allreduce_f32(self, buf, n, average, on_comm): pass
    def nccl_broadcast(self, buf, nbytes, root): pass
    def nccl_wait(self): pass
    def nccl_fork(self): pass
    def nccl_destroy(self): pass


def install():
    """Route the cuda backend's ctypes layer to a FakeDevice.  Returns (device, undo)."""
    dev = FakeDevice()
    saved = (rt.api, rt._ready, rt.load, rt.ensure_device)
    rt.api, rt._ready = dev, True
    rt.load = lambda: dev
    rt.ensure_device = lambda device=-1: dev

    def undo():
        rt.api, rt._ready, rt.load, rt.ensure_device = saved
    return dev, undo
