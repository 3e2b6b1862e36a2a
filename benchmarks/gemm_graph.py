"""Per-shape tensor-core GEMM time INSIDE a CUDA graph (no python / launch latency between kernels).

opbench's eager loop is bound by interpreter dispatch below ~30 us per call, which hides what the BERT
step's one-wave GEMMs really cost.  Here each shape is captured `reps` times back-to-back into one graph
(rotating over enough operand sets to exceed L2, like the step does) and the replay is timed with events.

    python benchmarks/gemm_graph.py [--reps 24]
"""
import argparse
import json
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightgrad_b200 import CudaTensor                                  # noqa: E402
from lightgrad_b200.autograd.cuda import ops, runtime as rt            # noqa: E402
from lightgrad_b200.autograd.cuda.graph import StepGraph               # noqa: E402
from lightgrad_b200.autograd.cuda.ops import _gemm, _gemm_grouped, _gemm_epilogue, _attention_gemm, _swap_last   # noqa: E402

R, H, F, V = 4096, 768, 3072, 30522          # rows = batch 32 x seq 128


def rand(*shape):
    return CudaTensor.from_numpy(np.random.uniform(-1, 1, shape).astype(np.float32))


def rand_pitched(rows, cols):
    """(rows, cols) view of a buffer whose row pitch is a multiple of 32 floats (what the framework's own GEMM
    results with an unaligned width look like, e.g. the 30522-wide logits)."""
    ld = (cols + 31) // 32 * 32
    full = rand(rows, ld)
    return full if ld == cols else full._view((rows, cols), (ld, 1))


def cases():
    # name, builder -> (callable issuing ONE launch on operand set i, flops per launch)
    def fwd(n_in, n_out):
        def make(sets):
            X = [rand(R, n_in) for _ in range(sets)]
            W = [rand(n_out, n_in) for _ in range(sets)]
            b = rand(n_out)
            O = [rand_pitched(R, n_out) for _ in range(sets)]
            return (lambda i: _gemm(X[i], _swap_last(W[i]), out=O[i], bias=b)), 2.0 * R * n_in * n_out
        return make

    def dx(n_in, n_out):
        def make(sets):
            G = [rand_pitched(R, n_out) for _ in range(sets)]
            W = [rand(n_out, n_in) for _ in range(sets)]
            O = [CudaTensor.empty((R, n_in)) for _ in range(sets)]
            return (lambda i: _gemm(G[i], W[i], out=O[i])), 2.0 * R * n_in * n_out
        return make

    def dw(n_in, n_out):
        def make(sets):
            G = [rand_pitched(R, n_out) for _ in range(sets)]
            X = [rand(R, n_in) for _ in range(sets)]
            O = [CudaTensor.zeros((n_out, n_in)) for _ in range(sets)]
            return (lambda i: _gemm(_swap_last(G[i]), X[i], out=O[i], accumulate=True)), 2.0 * R * n_in * n_out
        return make

    def qkv_fwd():
        def make(sets):
            X = [rand(R, H) for _ in range(sets)]
            W = [[rand(H, H) for _ in range(3)] for _ in range(sets)]
            b = [rand(H) for _ in range(3)]
            O = [CudaTensor.empty((3, R, H)) for _ in range(sets)]
            def go(i):
                parts = [O[i]._view((R, H), (H, 1), g * R * H) for g in range(3)]
                _gemm_grouped([X[i]] * 3, [_swap_last(w) for w in W[i]], parts, b)
            return go, 3 * 2.0 * R * H * H
        return make

    def qkv_dx():
        def make(sets):
            G = [rand(3, R, H) for _ in range(sets)]
            W = [[rand(H, H) for _ in range(3)] for _ in range(sets)]
            O = [CudaTensor.empty((R, H)) for _ in range(sets)]
            def go(i):
                parts = [G[i]._view((R, H), (H, 1), g * R * H) for g in range(3)]
                _gemm_grouped(parts, W[i], [O[i]] * 3)
            return go, 3 * 2.0 * R * H * H
        return make

    def qkv_dw():
        def make(sets):
            G = [rand(3, R, H) for _ in range(sets)]
            X = [rand(R, H) for _ in range(sets)]
            O = [[CudaTensor.zeros((H, H)) for _ in range(3)] for _ in range(sets)]
            def go(i):
                parts = [_swap_last(G[i]._view((R, H), (H, 1), g * R * H)) for g in range(3)]
                _gemm_grouped(parts, [X[i]] * 3, O[i], accumulate=True)
            return go, 3 * 2.0 * R * H * H
        return make

    def epi(kind):
        def make(sets):
            if kind == 1:      # h = x W1^T + b1, act = gelu(h)
                X = [rand(R, H) for _ in range(sets)]
                W = [rand(F, H) for _ in range(sets)]
                b = rand(F)
                O = [CudaTensor.empty((R, F)) for _ in range(sets)]
                A = [CudaTensor.empty((R, F)) for _ in range(sets)]
                return (lambda i: _gemm_epilogue(X[i], _swap_last(W[i]), O[i], b, 1, A[i])), 2.0 * R * H * F
            G = [rand(R, H) for _ in range(sets)]
            W = [rand(H, F) for _ in range(sets)]
            Hh = [rand(R, F) for _ in range(sets)]
            O = [CudaTensor.empty((R, F)) for _ in range(sets)]
            return (lambda i: _gemm_epilogue(G[i], W[i], O[i], None, 2, Hh[i])), 2.0 * R * H * F
        return make

    def att(kind):
        b, h, s, d = 32, 12, 128, 64
        def make(sets):
            Q = [rand(b, h, s, d) for _ in range(sets)]
            K = [rand(b, h, s, d) for _ in range(sets)]
            P = [rand(b, h, s, s) for _ in range(sets)]
            if kind == 'qk':
                return (lambda i: _gemm(Q[i], _swap_last(K[i]), out=P[i])), 2.0 * b * h * s * s * d
            if kind == 'qk_softmax':
                return (lambda i: _attention_gemm(Q[i], _swap_last(K[i]), P[i], 3, 0.125)), 2.0 * b * h * s * s * d
            if kind == 'dp_softmax_bwd':
                D = [CudaTensor.empty((b, h, s, s)) for _ in range(sets)]
                return (lambda i: _attention_gemm(Q[i], _swap_last(K[i]), D[i], 4, 0.125, aux=P[i])), \
                    2.0 * b * h * s * s * d
            O = [CudaTensor.empty((b, h, s, d)) for _ in range(sets)]
            if kind == 'pv':
                return (lambda i: _gemm(P[i], K[i], out=O[i])), 2.0 * b * h * s * s * d
            return (lambda i: _gemm(_swap_last(P[i]), K[i], out=O[i])), 2.0 * b * h * s * s * d
        return make

    def layout(M, N, K, a_mn, b_mn, bias):
        def make(sets):
            A = [rand(K, M) if a_mn else rand(M, K) for _ in range(sets)]
            B = [rand(K, N) if b_mn else rand(N, K) for _ in range(sets)]
            bv = rand(N) if bias else None
            O = [CudaTensor.empty((M, N)) for _ in range(sets)]
            return (lambda i: _gemm(_swap_last(A[i]) if a_mn else A[i], B[i] if b_mn else _swap_last(B[i]),
                                    out=O[i], bias=bv)), 2.0 * M * N * K
        return make

    if os.environ.get('GEMM_GRAPH_SUITE') == 'fused':
        # the non-matmul kernels of the BERT step at their step shapes; "flops" slot carries BYTES here
        import lightgrad_b200 as light
        RED, EW = rt.RED, rt.EW

        def ln(bwd):
            def make(sets):
                X = [rand(R, H) for _ in range(sets)]
                G = [rand(R, H) for _ in range(sets)]
                w, b = rand(H), rand(H)
                wg = CudaTensor.zeros((2, H))
                Y = [CudaTensor.empty((R, H)) for _ in range(sets)]
                mean, rstd = CudaTensor.empty((R,)), CudaTensor.empty((R,))
                rt.api.layernorm_fwd(rt.F32, X[0].ptr, w.ptr, b.ptr, Y[0].ptr, mean.ptr, rstd.ptr, R, H, 1e-5)
                if not bwd:
                    return (lambda i: rt.api.layernorm_fwd(rt.F32, X[i].ptr, w.ptr, b.ptr, Y[i].ptr, mean.ptr, rstd.ptr,
                                                           R, H, 1e-5)), 2.0 * R * H * 4
                return (lambda i: rt.api.layernorm_bwd(rt.F32, X[i].ptr, w.ptr, mean.ptr, rstd.ptr, G[i].ptr, Y[i].ptr,
                                                       wg.ptr, wg.ptr + 4 * H, R, H, 1, None)), 3.0 * R * H * 4
            return make

        def colsum(n):
            def make(sets):
                G = [rand(R, n) for _ in range(sets)]
                bg = CudaTensor.zeros((n,))
                return (lambda i: rt.api.reduce_pitched(RED['SUM'], rt.F32, G[i].ptr, bg.ptr, 1, R, n, n, 1.0, 1)), \
                    1.0 * R * n * 4
            return make

        def ew(op, n_in, n):
            def make(sets):
                A = [rand(R, n) for _ in range(sets)]
                B = [rand(R, n) for _ in range(sets)]
                O = [CudaTensor.empty((R, n)) for _ in range(sets)]
                return (lambda i: rt.api.ew_flat(EW[op], rt.F32, A[i].ptr, B[i].ptr if n_in > 1 else None, None,
                                                 O[i].ptr, R * n, 0.0)), (n_in + 1.0) * R * n * 4
            return make

        def softmax(bwd):
            rows, cols = 32 * 12 * 128, 128
            def make(sets):
                Xs = [rand(rows, cols) for _ in range(sets)]
                Gs = [rand(rows, cols) for _ in range(sets)]
                Ys = [CudaTensor.empty((rows, cols)) for _ in range(sets)]
                if not bwd:
                    return (lambda i: rt.api.softmax_fwd(rt.F32, Xs[i].ptr, Ys[i].ptr, rows, cols, 0.125)), \
                        2.0 * rows * cols * 4
                return (lambda i: rt.api.softmax_bwd(rt.F32, Xs[i].ptr, Gs[i].ptr, Ys[i].ptr, rows, cols, 0.125)), \
                    3.0 * rows * cols * 4
            return make

        def ce(bwd):
            ld = (V + 31) // 32 * 32
            def make(sets):
                X = [rand(R, ld) for _ in range(sets)]
                lab = CudaTensor.from_numpy(np.random.randint(0, V, size=(R,)).astype(np.int32), requires_grad=False)
                loss_rows, lse, one = CudaTensor.empty((R,)), CudaTensor.empty((R,)), CudaTensor.ones((1,))
                D = [CudaTensor.empty((R, ld)) for _ in range(sets)]
                rt.api.cross_entropy_fwd(rt.F32, rt.I32, X[0].ptr, ld, lab.ptr, loss_rows.ptr, lse.ptr, R, V)
                if not bwd:
                    return (lambda i: rt.api.cross_entropy_fwd(rt.F32, rt.I32, X[i].ptr, ld, lab.ptr, loss_rows.ptr,
                                                               lse.ptr, R, V)), 1.0 * R * V * 4
                return (lambda i: rt.api.cross_entropy_bwd(rt.F32, rt.I32, X[i].ptr, ld, lab.ptr, lse.ptr, one.ptr,
                                                           D[i].ptr, ld, R, V)), 2.0 * R * V * 4
            return make

        def adam():
            def make(sets):
                n = 110 * 1000 * 1000 // 64 * 64
                ps = [CudaTensor.zeros((n,))]
                opt = light.optim.Adam(ps, lr=1e-4)
                ps[0].zero_grad()
                opt.step()
                return (lambda i: opt.step()), 7.0 * n * 4
            return make

        def attention(bwd):
            def make(sets):
                B, S, NH, DH = 32, 128, 12, 64
                QKV = [rand(3, R, H) for _ in range(sets)]
                O = [CudaTensor.empty((R, H)) for _ in range(sets)]
                DO = [rand(R, H) for _ in range(sets)]
                DQKV = [CudaTensor.empty((3, R, H)) for _ in range(sets)]
                lse = CudaTensor.empty((B * NH * S,))
                db = CudaTensor.zeros((3, H))
                scale = 1.0 / 8.0
                for i in range(sets):
                    rt.api.attention_fwd(rt.F32, QKV[i].ptr, B, S, NH, DH, scale, O[i].ptr, lse.ptr)
                if not bwd:
                    # algorithmic bytes: Q, K, V read + O written
                    return (lambda i: rt.api.attention_fwd(rt.F32, QKV[i].ptr, B, S, NH, DH, scale, O[i].ptr,
                                                           lse.ptr)), 4.0 * R * H * 4
                # Q, K, V, O, dO read + dQ, dK, dV written
                return (lambda i: rt.api.attention_bwd(rt.F32, QKV[i].ptr, O[i].ptr, DO[i].ptr, lse.ptr, B, S, NH, DH,
                                                       scale, DQKV[i].ptr, db.ptr, db.ptr + 4 * H,
                                                       db.ptr + 8 * H)), 8.0 * R * H * 4
            return make

        return [('attention fwd 32x12x128x64', attention(False)), ('attention bwd 32x12x128x64', attention(True)),
                ('layernorm fwd 4096x768', ln(False)), ('layernorm bwd 4096x768 (+partials reduce)', ln(True)),
                ('bias grad colsum 4096x768', colsum(H)), ('bias grad colsum 4096x3072', colsum(F)),
                ('add 4096x768', ew('ADD', 2, H)), ('gelu 4096x3072', ew('GELU', 1, F)),
                ('gelu_bwd 4096x3072', ew('GELU_BWD', 2, F)),
                ('softmax fwd 49152x128', softmax(False)), ('softmax bwd 49152x128', softmax(True)),
                ('cross entropy fwd 4096x30522', ce(False)), ('cross entropy bwd 4096x30522', ce(True)),
                ('adam 110M', adam())]

    if os.environ.get('GEMM_GRAPH_SUITE') == 'layout':
        out = []
        for (M, N, K) in [(4096, 3072, 768), (4096, 768, 3072), (4096, 768, 768), (4096, 4096, 4096)]:
            for a_mn in (0, 1):
                for b_mn in (0, 1):
                    for bias in ((0, 1) if (a_mn, b_mn) == (0, 0) else (0,)):
                        out.append(('%dx%dx%d A_%s B_%s%s' % (M, N, K, 'mn' if a_mn else 'k', 'mn' if b_mn else 'k',
                                                            ' +bias' if bias else ''),
                                    layout(M, N, K, a_mn, b_mn, bias)))
        return out

    return [
        ('proj fwd 4096x768x768', fwd(H, H)), ('proj dX', dx(H, H)), ('proj dW (acc)', dw(H, H)),
        ('ffn1 fwd 4096x3072x768', fwd(H, F)), ('ffn1 dX', dx(H, F)), ('ffn1 dW (acc)', dw(H, F)),
        ('ffn2 fwd 4096x768x3072', fwd(F, H)), ('ffn2 dX', dx(F, H)), ('ffn2 dW (acc)', dw(F, H)),
        ('ffn1 fwd + gelu epilogue (2 results)', epi(1)), ('ffn2 dX * gelu\' epilogue', epi(2)),
        ('qkv grouped fwd', qkv_fwd()), ('qkv k-concat dX', qkv_dx()), ('qkv grouped dW (acc)', qkv_dw()),
        ('attn QK^T 384x128x128x64', att('qk')), ('attn QK^T + softmax epilogue', att('qk_softmax')),
        ('attn dO V^T + softmax-bwd epilogue', att('dp_softmax_bwd')), ('attn PV', att('pv')), ('attn P^T dO', att('ptv')),
        ('decoder fwd 4096x30522x768', fwd(H, V)), ('decoder dX', dx(H, V)), ('decoder dW (acc)', dw(H, V)),
    ]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=24)
    ap.add_argument('--sets', type=int, default=6)
    ap.add_argument('--only', default='')
    a = ap.parse_args()
    api = rt.ensure_device()
    ops.set_matmul_mode('tf32')
    for name, make in cases():
        if a.only and a.only not in name:
            continue
        sets = 2 if ('decoder' in name or '4096x4096' in name or 'cross entropy' in name) else (12 if '768' in name and 'x3072' not in name and os.environ.get('GEMM_GRAPH_SUITE') == 'fused' else a.sets)
        go, flops = make(sets)

        def body():
            for r in range(a.reps):
                go(r % sets)
        g = StepGraph(body, warmup=1)
        for _ in range(2):
            g.replay()
        e0, e1 = rt.Event(), rt.Event()
        n = 5
        e0.record()
        for _ in range(n):
            g.replay()
        e1.record()
        e1.synchronize()
        us = e0.elapsed_ms(e1) * 1e3 / (n * a.reps)
        unit = 'gbps' if os.environ.get('GEMM_GRAPH_SUITE') == 'fused' else 'tflops'
        print(json.dumps({'case': name, 'us_per_launch': round(us, 2), unit: round(flops / us * (1e-3 if unit == 'gbps' else 1e-6), 1),
                          'launches_in_graph': g.n_kernels}), flush=True)
        del g, go
        api.empty_cache()


if __name__ == '__main__':
    main()
