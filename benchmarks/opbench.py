"""Op micro-benchmarks (BASELINE.json configs 2 and 3): elementwise / reduction HBM GB/s and matmul
TFLOP/s on one B200, CUDA-event timed, printed as JSON lines.

    python benchmarks/opbench.py [--suite ew,reduce,gemm] [--max-log2 28] [--cpu]

Algorithmic bytes per element are the SURVEY.md 8(d) figures and are printed with every line.
`--cpu` also times the numpy oracle (single-threaded elementwise, OpenBLAS GEMM) on the host cores.
"""
import argparse
import json
import os
import sys
import time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

PEAKS = {'hbm_gbs': 6550.7, 'bf16_tflops': 1636.6}
try:
    PEAKS.update(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                             'MEASURED_PEAKS.json'))))
except Exception:
    pass


def time_gpu(fn, iters, warmup=3, flush=None):
    """Average device time of fn(): `iters` launches queued back to back between two CUDA events on the
    compute stream (so host dispatch latency is not measured).  When `flush` is given it is queued before
    every launch (evicts L2) and its own time, measured the same way, is subtracted."""
    from lightgrad_b200.autograd.cuda import runtime as rt

    def span(body):
        rt.synchronize()
        e0 = rt.Event().record()
        for _ in range(iters):
            body()
        e1 = rt.Event().record()
        e1.synchronize()
        return e0.elapsed_ms(e1)
    for _ in range(warmup):
        fn()
    if flush is None:
        return span(fn) / iters
    for _ in range(2):
        flush()
    t_flush = span(flush)

    def both():
        flush()
        fn()
    return max(span(both) - t_flush, 1e-6) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--suite', default='ew,reduce,gemm')
    ap.add_argument('--max-log2', type=int, default=28)
    ap.add_argument('--min-log2', type=int, default=10)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--cpu', action='store_true')
    ap.add_argument('--modes', default='fp32')
    args = ap.parse_args()
    import lightgrad_b200 as light
    from lightgrad_b200 import CudaTensor as T
    from lightgrad_b200.autograd.cuda import runtime as rt, ops
    suites = args.suite.split(',')
    # a 256 MiB scratch write between timed launches evicts the 126 MB L2
    scratch = T.empty((64 << 20,), requires_grad=False)
    flush = lambda: scratch.fill(0.0)  # noqa: E731
    rs = np.random.RandomState(0)

    def emit(**kw):
        print(json.dumps(kw), flush=True)

    if 'ew' in suites:
        for lg in range(args.min_log2, args.max_log2 + 1, 2):
            n = 1 << lg
            a = T.from_numpy(rs.uniform(-1, 1, n).astype(np.float32))
            b = T.from_numpy(rs.uniform(-1, 1, n).astype(np.float32))
            g = T.from_numpy(rs.uniform(-1, 1, n).astype(np.float32))
            with light.no_grad():
                cases = [('add', 12, lambda: a + b), ('mul', 12, lambda: a * b), ('relu', 8, lambda: a.relu()),
                         ('exp', 8, lambda: a.exp()), ('gelu', 8, lambda: a.gelu()),
                         ('relu_bwd', 12, lambda: ops._ewn(rt.EW['RELU_BWD'], (a, g))),
                         ('mul_bwd', 20, lambda: rt.api.ew_bwd2_flat(0, rt.F32, a.ptr, b.ptr, g.ptr, a2.ptr, b2.ptr, n))]
                a2, b2 = T.empty((n,)), T.empty((n,))
                for name, bpe, fn in cases:
                    ms = time_gpu(fn, args.iters, flush=flush if n * 4 < (256 << 20) else None)
                    gbs = bpe * n / ms / 1e6
                    emit(suite='ew', op=name, log2n=lg, ms=round(ms, 5), bytes_per_elem=bpe, gbs=round(gbs, 1),
                         frac_of_measured_hbm=round(gbs / PEAKS['hbm_gbs'], 3))
                if args.cpu and lg <= 24:
                    an, bn = a.numpy(), b.numpy()
                    for name, bpe, fn in (('add', 12, lambda: an + bn), ('exp', 8, lambda: np.exp(an))):
                        t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
                        emit(suite='ew', op=name, log2n=lg, impl='cpu-numpy', ms=round(dt * 1e3, 3),
                             gbs=round(bpe * n / dt / 1e9, 2), cores=1)
            # broadcast add (R, C) + (C,)
            R = n // 1024
            m = a.reshape(R, 1024)
            bias = T.from_numpy(rs.uniform(-1, 1, 1024).astype(np.float32))
            with light.no_grad():
                ms = time_gpu(lambda: m + bias, args.iters, flush=flush if n * 4 < (256 << 20) else None)
            emit(suite='ew', op='add_rowbcast', log2n=lg, ms=round(ms, 5), bytes_per_elem=8,
                 gbs=round(8 * n / ms / 1e6, 1), frac_of_measured_hbm=round(8 * n / ms / 1e6 / PEAKS['hbm_gbs'], 3))
            del a, b, g, a2, b2, m

    if 'reduce' in suites:
        for lg in range(args.min_log2, args.max_log2 + 1, 2):
            n = 1 << lg
            side = 1 << (lg // 2)
            x = T.from_numpy(rs.uniform(-1, 1, (side, n // side)).astype(np.float32))
            with light.no_grad():
                for name in ('sum', 'max'):
                    for axis in (None, 0, 1):
                        fn = (lambda nm=name, ax=axis: getattr(x, nm)(axis=ax))
                        ms = time_gpu(fn, args.iters, flush=flush if n * 4 < (256 << 20) else None)
                        gbs = 4 * n / ms / 1e6
                        emit(suite='reduce', op=name, axis=axis, log2n=lg, ms=round(ms, 5), bytes_per_elem=4,
                             gbs=round(gbs, 1), frac_of_measured_hbm=round(gbs / PEAKS['hbm_gbs'], 3))
            del x

    if 'gemm' in suites:
        shapes = [(s, s, s) for s in (256, 512, 1024, 2048, 4096, 8192)] + \
                 [(4096, 768, 768), (4096, 3072, 768), (4096, 768, 3072), (4096, 30522, 768), (64, 128, 784)]
        for mode in args.modes.split(','):
            ops.set_matmul_mode(mode)
            for (M, N, K) in shapes:
                if mode == 'fp32' and M * N * K > 4096 ** 3:
                    continue
                a = T.from_numpy(rs.uniform(-1, 1, (M, K)).astype(np.float32))
                w = T.from_numpy(rs.uniform(-1, 1, (N, K)).astype(np.float32))
                with light.no_grad():
                    ms = time_gpu(lambda: a.linear(w), max(3, args.iters // 2))
                tf = 2.0 * M * N * K / ms / 1e9
                peak = 74.0 if mode == 'fp32' else (PEAKS['bf16_tflops'] / 2 if mode == 'tf32' else PEAKS['bf16_tflops'])
                emit(suite='gemm', mode=mode, M=M, N=N, K=K, ms=round(ms, 4), tflops=round(tf, 2),
                     frac_of_peak=round(tf / peak, 3), peak_tflops=peak)
                if args.cpu and M * N * K <= 4096 ** 3 and mode == args.modes.split(',')[0]:
                    an, wn = a.numpy(), w.numpy()
                    t0 = time.perf_counter(); an @ wn.T; dt = time.perf_counter() - t0
                    emit(suite='gemm', impl='cpu-openblas', M=M, N=N, K=K, ms=round(dt * 1e3, 2),
                         tflops=round(2.0 * M * N * K / dt / 1e12, 3), cores=os.cpu_count())
                del a, w
            # batched attention shapes of BASELINE config 3: (384, 128, 64) x (384, 64, 128) and P @ V
            for (Bt, M, K, N) in ((384, 128, 64, 128), (384, 128, 128, 64)):
                a = T.from_numpy(rs.uniform(-1, 1, (Bt, M, K)).astype(np.float32))
                b = T.from_numpy(rs.uniform(-1, 1, (Bt, K, N)).astype(np.float32))
                with light.no_grad():
                    ms = time_gpu(lambda: ops._gemm(a, b), max(3, args.iters // 2))
                tf = 2.0 * Bt * M * N * K / ms / 1e9
                emit(suite='gemm', mode=mode, batch=Bt, M=M, N=N, K=K, ms=round(ms, 4), tflops=round(tf, 2),
                     note='host-dispatch bound at this size; see gemm_fit.py for device time')
                del a, b
            ops.set_matmul_mode('fp32')


    if 'gemm_bwd' in suites:
        # BASELINE config 3: forward + both backward GEMMs (dA = dC.B^T, dB = A^T.dC), 6*M*N*K flops
        shapes = [(s_, s_, s_) for s_ in (1024, 2048, 4096)] + [(4096, 768, 768), (4096, 3072, 768), (4096, 768, 3072)]
        for mode in args.modes.split(','):
            ops.set_matmul_mode(mode)
            for (M, N, K) in shapes:
                a = T.from_numpy(rs.uniform(-1, 1, (M, K)).astype(np.float32))
                b = T.from_numpy(rs.uniform(-1, 1, (K, N)).astype(np.float32))
                g = T.from_numpy(rs.uniform(-1, 1, (M, N)).astype(np.float32))

                def fwd_bwd():
                    with light.no_grad():
                        ops._gemm(a, b)
                        ops._gemm(g, ops._swap_last(b))
                        ops._gemm(ops._swap_last(a), g)
                ms = time_gpu(fwd_bwd, max(3, args.iters // 2))
                tf = 6.0 * M * N * K / ms / 1e9
                peak = 74.0 if mode == 'fp32' else PEAKS['bf16_tflops'] / 2
                emit(suite='gemm_bwd', mode=mode, M=M, N=N, K=K, ms=round(ms, 4), tflops=round(tf, 2),
                     frac_of_peak=round(tf / peak, 3), peak_tflops=peak, flops='6*M*N*K')
                if args.cpu and mode == args.modes.split(',')[0]:
                    an, bn, gn = a.numpy(), b.numpy(), g.numpy()
                    t0 = time.perf_counter(); an @ bn; gn @ bn.T; an.T @ gn; dt = time.perf_counter() - t0
                    emit(suite='gemm_bwd', impl='cpu-openblas', M=M, N=N, K=K, ms=round(dt * 1e3, 2),
                         tflops=round(6.0 * M * N * K / dt / 1e12, 3), cores=os.cpu_count())
                del a, b, g
        ops.set_matmul_mode('fp32')

    if 'bwd' in suites:
        # reduction / unary backward passes of BASELINE config 2 (through the autograd API, incl. the walk)
        for lg in range(max(args.min_log2, 20), args.max_log2 + 1, 2):
            n = 1 << lg
            side = 1 << (lg // 2)
            x = T.from_numpy(rs.uniform(-1, 1, (side, n // side)).astype(np.float32))
            for name, bpe in (('sum', 4), ('max', 8), ('exp', 12), ('relu', 12)):
                y = getattr(x, name)()
                gy = T.ones(y.shape, requires_grad=False)

                def bwd():
                    with light.no_grad():
                        return y.ctx.backward(gy)
                ms = time_gpu(bwd, args.iters)
                if name == 'sum':
                    # the backward of sum is a zero-stride view; the bytes move when it is accumulated
                    def bwd():  # noqa: F811
                        with light.no_grad():
                            return y.ctx.backward(gy).copy()
                    ms = time_gpu(bwd, args.iters)
                emit(suite='bwd', op=name + '_bwd', log2n=lg, ms=round(ms, 5), bytes_per_elem=bpe,
                     gbs=round(bpe * n / ms / 1e6, 1), frac_of_measured_hbm=round(bpe * n / ms / 1e6 / PEAKS['hbm_gbs'], 3))
            del x

    if 'mnist' in suites:
        # BASELINE config 1: MLP 784-128-10, batch 64, mse + SGD(lr=1e-4): launch-latency bound
        import lightgrad_b200.nn as nn
        from examples import mnist as mn
        from lightgrad_b200.autograd.cuda.graph import StepGraph
        np.random.seed(0)
        model = mn.NN()
        opt = light.optim.SGD(model.parameters(), lr=1e-4)
        xb, yb = mn.synthetic_batch(64, seed=0)
        xd, yd = T.from_numpy(xb, requires_grad=False), T.from_numpy(yb, requires_grad=False)
        one_hot = T.zeros((64, 10), requires_grad=False)
        one_hot[range(64), yd] = 1

        def step():
            y = model(xd)
            loss = light.loss.mse(y, one_hot)
            opt.zero_grad()
            loss.backward()
            opt.step()
            return loss
        for _ in range(5):
            step()
        rt.synchronize()
        n0 = rt.launch_count()
        t0 = time.perf_counter()
        for _ in range(200):
            step()
        rt.synchronize()
        dt = (time.perf_counter() - t0) / 200
        emit(suite='mnist', impl='cuda-eager', ms_per_step=round(dt * 1e3, 4), samples_per_s=round(64 / dt, 1),
             launches_per_step=(rt.launch_count() - n0) / 200)
        sg = StepGraph(step, warmup=0)
        for _ in range(5):
            sg.replay()
        rt.synchronize()
        t0 = time.perf_counter()
        for _ in range(1000):
            sg.replay()
        rt.synchronize()
        dt = (time.perf_counter() - t0) / 1000
        emit(suite='mnist', impl='cuda-graph', ms_per_step=round(dt * 1e3, 4), samples_per_s=round(64 / dt, 1),
             kernels_per_step=sg.n_kernels)
        if args.cpu:
            from oracle import CpuTensor
            with nn.use_tensor(CpuTensor):
                np.random.seed(0)
                cm = mn.NN()
            copt = light.optim.SGD(cm.parameters(), lr=1e-4)
            cx, cy = CpuTensor.from_numpy(xb, requires_grad=False), CpuTensor.from_numpy(yb, requires_grad=False)
            for _ in range(5):
                mn.train_step(cm, copt, cx, cy, CpuTensor)
            t0 = time.perf_counter()
            for _ in range(200):
                mn.train_step(cm, copt, cx, cy, CpuTensor)
            dt = (time.perf_counter() - t0) / 200
            emit(suite='mnist', impl='cpu-oracle', ms_per_step=round(dt * 1e3, 4), samples_per_s=round(64 / dt, 1),
                 cores=os.cpu_count())


if __name__ == '__main__':
    main()
