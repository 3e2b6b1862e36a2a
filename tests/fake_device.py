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
      rt.U8: np.uint8, rt.I8: np.int8, rt.BF16: np.uint16}     # bf16 staging copies: raw 16-bit patterns


def _to_bf16(f):
    """float32 -> bf16 bit patterns, round to nearest even (what lg_cast does)."""
    u = np.ascontiguousarray(f, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = u + 0x7FFF + ((u >> 16) & 1)
    return (u >> 16).astype(np.uint16)


def _from_bf16(h):
    return (np.asarray(h).astype(np.uint32) << 16).view(np.float32)


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
        x = _arr(src, sd, shape, _lst(ss, nd))
        if dd == rt.BF16:
            assert sd == rt.F32
            _arr(dst, dd, shape, _lst(ds, nd))[...] = _to_bf16(x).reshape(x.shape)
            return
        if sd == rt.BF16:
            x = _from_bf16(x)
        _arr(dst, dd, shape, _lst(ds, nd))[...] = x.astype(NP[dd])

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
        if dt == rt.BF16:
            # bf16 operands (staging copies), fp32 accumulation and result -- only in the bf16 tensor-core mode
            assert mode == rt.GEMM_BF16_TC and self.gemm_tc_supported(mode, dt, dref)
            A, B, dt = _from_bf16(A), _from_bf16(B), rt.F32
        else:
            assert mode != rt.GEMM_BF16_TC, "bf16 mode takes bf16 operands"
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
        self.gemm(mode, dt, dref, a, b, c, None if epi == 2 else bias, 0)
        if dt == rt.BF16:
            dt = rt.F32
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
            if bias:            # for this epilogue `bias` names where the result's column sums are accumulated
                _arr(bias, dt, [d.N])[...] += Cm.sum(axis=0)

    def gemm_sm_limit(self, n): pass

    def comm_compute_begin(self): pass

    def comm_compute_end(self): pass

    def side_begin(self): pass

    def side_end(self): pass

    def side_join(self): pass

    def gemm_tc_supported(self, mode, dt, dref):
        if mode == rt.GEMM_TF32_TC:
            return 1 if dt == rt.F32 else 0
        if mode != rt.GEMM_BF16_TC or dt not in (rt.F32, rt.BF16):
            return 0
        # the library's rules for the bf16 kernel: operand strides 16-byte multiples of bf16, not tiny
        d = dref._obj
        q = 8
        a_ok = (d.sa_k == 1 and (d.sa_m % q == 0 or d.M == 1)) or (d.sa_m == 1 and (d.sa_k % q == 0 or d.K == 1))
        b_ok = (d.sb_k == 1 and (d.sb_n % q == 0 or d.N == 1)) or (d.sb_n == 1 and (d.sb_k % q == 0 or d.K == 1))
        c_ok = d.sc_n == 1 and (d.sc_m % 4 == 0 or d.M == 1)
        for n, st, m in ((d.batch1, d.sa_b1, q), (d.batch0, d.sa_b0, q), (d.batch1, d.sb_b1, q), (d.batch0, d.sb_b0, q),
                         (d.batch1, d.sc_b1, 4), (d.batch0, d.sc_b0, 4)):
            if n > 1 and (st <= 0 or st % m):
                return 0
        big = d.M * d.N * d.K * d.batch0 * d.batch1 >= 64 * 64 * 64 * 8
        return 1 if (a_ok and b_ok and c_ok and big) else 0

    def prof_gemm(self, enable): pass

    # ---- fused attention: same contract as the kernels (fp32, seq 128, head_dim 64)
    def attention_supported(self, dt, seq, hd):
        return 1 if (dt == rt.F32 and seq == 128 and hd == 64) else 0

    @staticmethod
    def _heads(x, b, s, h, d):
        return x.reshape(b, s, h, d).transpose(0, 2, 1, 3)

    def attention_fwd(self, dt, qkv, b, s, h, d, scale, out, lse):
        self.launches += 1
        X = _arr(qkv, dt, [3, b * s, h * d])
        q, k, v = (self._heads(X[i], b, s, h, d) for i in range(3))
        sc = np.float32(scale) * (q @ k.transpose(0, 1, 3, 2))
        m = sc.max(axis=-1, keepdims=True)
        e = np.exp(sc - m)
        z = e.sum(axis=-1, keepdims=True)
        o = (e / z) @ v
        _arr(out, dt, [b, s, h, d])[...] = o.transpose(0, 2, 1, 3)
        _arr(lse, dt, [b, h, s])[...] = (m + np.log(z))[..., 0]

    def attention_bwd(self, dt, qkv, out, dout, lse, b, s, h, d, scale, dqkv, dbq, dbk, dbv):
        self.launches += 1
        X = _arr(qkv, dt, [3, b * s, h * d])
        q, k, v = (self._heads(X[i], b, s, h, d) for i in range(3))
        o = self._heads(_arr(out, dt, [b * s, h * d]), b, s, h, d)
        do = self._heads(_arr(dout, dt, [b * s, h * d]), b, s, h, d)
        L = _arr(lse, dt, [b, h, s])[..., None]
        p = np.exp(np.float32(scale) * (q @ k.transpose(0, 1, 3, 2)) - L)
        dp = do @ v.transpose(0, 1, 3, 2)
        ds = np.float32(scale) * p * (dp - (do * o).sum(axis=-1, keepdims=True))
        G = _arr(dqkv, dt, [3, b, s, h, d])
        G[0] = (ds @ k).transpose(0, 2, 1, 3)
        G[1] = (ds.transpose(0, 1, 3, 2) @ q).transpose(0, 2, 1, 3)
        G[2] = (p.transpose(0, 1, 3, 2) @ do).transpose(0, 2, 1, 3)
        for i, db in enumerate((dbq, dbk, dbv)):
            if db:
                _arr(db, dt, [h * d])[...] += G[i].reshape(b * s, h * d).sum(axis=0)

    def prof_gemm_read(self, ms, n, fl):
        ms._obj.value, n._obj.value, fl._obj.value = 0.0, 0, 0.0

    # ---- convolution lowering
    @staticmethod
    def _windows(x, n, k, s):
        shape = x.shape[:-n] + tuple((d - kk) // ss + 1 for d, kk, ss in zip(x.shape[-n:], k, s)) + tuple(k)
        strides = x.strides[:-n] + tuple(ts * ss for ts, ss in zip(x.strides[-n:], s)) + x.strides[-n:]
        return np.lib.stride_tricks.as_strided(x, shape=shape, strides=strides)

    def im2col(self, dt, n, lead, in_dims, k_dims, strides, x, cols):
        self.launches += 1
        ind, k, s = _lst(in_dims, n), _lst(k_dims, n), _lst(strides, n)
        X = _arr(x, dt, [lead] + ind)
        win = self._windows(X, n, k, s)
        _arr(cols, dt, win.shape)[...] = win

    def col2im(self, dt, n, lead, in_dims, k_dims, strides, cols, dx):
        self.launches += 1
        ind, k, s = _lst(in_dims, n), _lst(k_dims, n), _lst(strides, n)
        DX = _arr(dx, dt, [lead] + ind)
        DX[...] = 0
        win = self._windows(DX, n, k, s)
        Cm = _arr(cols, dt, win.shape)
        for kidx in np.ndindex(*k):
            sel = (slice(None),) * (1 + n) + tuple(kidx)
            win[sel] += Cm[sel]          # one kernel offset at a time: the windows of one offset never overlap

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

    def layernorm_bwd(self, dt, x, w, mean, rstd, g, dx, dw, db, rows, cols, accumulate, dx_colsum=None):
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
        if dx_colsum:
            _arr(dx_colsum, dt, [cols])[...] = _arr(dx, dt, [rows, cols]).sum(axis=0)

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

    # ---- NVLink multicast exchange: a team of ONE device (both mappings are the same host block; the reduce-scatter
    #      and all-gather are identities), which is what the host logic needs: region layout, arena migration, bucket
    #      ranges and optimizer arguments
    def mc_supported(self, yes):
        yes._obj.value = 1

    def mc_region_bytes(self, arena_bytes, world, region, goff, poff, foff):
        gran = 1 << 16
        a = (int(arena_bytes) + gran - 1) // gran * gran
        goff._obj.value, poff._obj.value, foff._obj.value, region._obj.value = 0, a, 2 * a, 2 * a + gran

    def mc_create(self, nbytes, world, fd):
        assert world == 1, "the test double has one device"
        self._mc_bytes, fd._obj.value = int(nbytes), -1

    def mc_import(self, fd, nbytes, world):
        raise AssertionError("the test double has one device")

    def mc_add_device(self): pass

    def mc_bind(self, local_ref, mc_ref):
        buf = np.zeros(self._mc_bytes + 256, dtype=np.uint8)
        addr = (buf.ctypes.data + 255) // 256 * 256
        self._mc_block, self._mc_base = buf, addr
        local_ref._obj.value = mc_ref._obj.value = addr

    def mc_release(self):
        self._mc_block = None

    def bucket_step(self, kind, p, g, m, v, lo, hi, n_seg, seg_end, t_dev, lr, b1, b2, eps, momentum, seg_offset,
                    t_advance):
        assert lo % 4 == 0 and hi % 4 == 0
        n = hi - lo
        if n == 0:
            return
        if kind <= 1:
            self.adam_step(kind, int(p) + lo * 4, int(g) + lo * 4, int(m) + lo * 4, int(v) + lo * 4, n, n_seg, seg_end,
                           t_dev, lr, b1, b2, eps, lo, seg_offset, t_advance)
        else:
            self.sgd_step(int(p) + lo * 4, int(g) + lo * 4, (int(m) + lo * 4) if (m and momentum != 0.0) else None, n, lr,
                          momentum)

    def mc_trace_mark(self): pass

    def mc_trace_read(self, out, max_records, n, reset):
        n._obj.value = 0

    def mc_exchange_step(self, kind, goff, poff, foff, lo, hi, rank, world, m, v, n_seg, seg_end, t_dev, lr, b1, b2,
                         eps, momentum, seg_offset, t_advance):
        assert world == 1 and rank == 0 and lo % 4 == 0 and hi % 4 == 0
        n = hi - lo
        if n == 0:
            return
        g = self._mc_base + goff + lo * 4
        p = self._mc_base + poff + lo * 4
        if kind <= 1:
            self.adam_step(kind, p, g, int(m) + lo * 4, int(v) + lo * 4, n, n_seg, seg_end, t_dev, lr, b1, b2, eps, lo,
                           seg_offset, t_advance)
        elif kind == 2:
            self.sgd_step(p, g, (int(m) + lo * 4) if (m and momentum != 0.0) else None, n, lr, momentum)
        else:
            self.launches += 1


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
