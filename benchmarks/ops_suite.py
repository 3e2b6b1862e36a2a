"""BASELINE.json configs 1-3 as one bounded measurement, for bench.py's `ops` object.

    config 1   examples/mnist.py MLP 784-128-10, batch 64, mse + SGD: steps/s (CUDA-graph replay and eager)
    config 2   add / mul / relu / exp forward, relu / mul backward, sum / max (axis None, 0, 1) at 2^26 and 2^28
               fp32 elements: HBM GB/s with the SURVEY.md 8(d) bytes-per-element figure printed beside it
    config 3   matmul 4096^3 and 8192^3, forward and forward + both backward GEMMs, in every matmul mode

Device numbers: CUDA events on the compute stream around launches queued back to back (inputs of
>= 256 MB per operand exceed the 126 MB L2).  Every case is timed next to the reference CPU tensor's
arithmetic (numpy; OpenBLAS for matmul) on the host cores, on the same shapes where that takes
under a few seconds and on a stated smaller shape otherwise.
"""
import os
import time
import numpy as np

NOMINAL_TF = {'fp32': 148 * 128 * 2 * 1.965e9 / 1e12, 'tf32': 1125.0, 'bf16': 2250.0}


def _span(rt, body, iters):
    rt.synchronize()
    e0 = rt.Event().record()
    for _ in range(iters):
        body()
    e1 = rt.Event().record()
    e1.synchronize()
    return e0.elapsed_ms(e1) / iters


def _time(rt, fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    return _span(rt, fn, iters)


def _cpu_time(fn, reps=2):
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return best


def elementwise(rt, light, T, ops, peaks, cpu, log2s=(26, 28)):
    rs = np.random.RandomState(0)
    out = []
    hbm = peaks['hbm_gbs']
    for lg in log2s:
        n = 1 << lg
        side = 1 << (lg // 2)
        an = rs.uniform(-1, 1, n).astype(np.float32)
        bn = rs.uniform(-1, 1, n).astype(np.float32)
        a, b, g = T.from_numpy(an), T.from_numpy(bn), T.from_numpy(bn[::-1].copy())
        da, db = T.empty((n,)), T.empty((n,))
        x2 = a.reshape(side, n // side)
        with light.no_grad():
            cases = [('add', 12, lambda: a + b), ('mul', 12, lambda: a * b), ('relu', 8, lambda: a.relu()),
                     ('exp', 8, lambda: a.exp()),
                     ('relu_bwd', 12, lambda: ops._ewn(rt.EW['RELU_BWD'], (a, g))),
                     ('mul_bwd', 20, lambda: rt.api.ew_bwd2_flat(0, rt.F32, a.ptr, b.ptr, g.ptr, da.ptr, db.ptr, n)),
                     ('add_rowbcast', 8, lambda: x2 + x2[0])]
            for name in ('sum', 'max'):
                for axis in (None, 0, 1):
                    cases.append(('%s_axis_%s' % (name, axis), 4, (lambda nm=name, ax=axis: getattr(x2, nm)(axis=ax))))
            for name, bpe, fn in cases:
                ms = _time(rt, fn)
                gbs = bpe * n / ms / 1e6
                out.append({'op': name, 'log2n': lg, 'ms': round(ms, 4), 'bytes_per_elem': bpe, 'gbs': round(gbs, 1),
                            'frac_of_measured_hbm': round(gbs / hbm, 3), 'frac_of_8tbs': round(gbs / 8000.0, 3)})
        if cpu and lg == log2s[0]:
            x2n = an.reshape(side, n // side)
            for name, bpe, fn in (('add', 12, lambda: an + bn), ('mul', 12, lambda: an * bn),
                                  ('relu', 8, lambda: np.maximum(an, 0)), ('exp', 8, lambda: np.exp(an)),
                                  ('sum_axis_None', 4, lambda: x2n.sum()), ('sum_axis_0', 4, lambda: x2n.sum(axis=0)),
                                  ('max_axis_1', 4, lambda: x2n.max(axis=1))):
                dt = _cpu_time(fn)
                out.append({'op': name, 'log2n': lg, 'impl': 'cpu-numpy (reference CpuTensor arithmetic)',
                            'ms': round(dt * 1e3, 2), 'gbs': round(bpe * n / dt / 1e9, 2), 'cores': 1})
        del a, b, g, da, db, x2
    return out


def matmul(rt, light, T, ops, peaks, cpu, modes, sizes=(4096, 8192)):
    rs = np.random.RandomState(1)
    out = []
    measured = {'fp32': None, 'tf32': peaks['bf16_tflops'] / 2.0, 'bf16': peaks['bf16_tflops']}
    prev = ops.get_matmul_mode()
    try:
        for s in sizes:
            an = rs.uniform(-1, 1, (s, s)).astype(np.float32)
            bn = rs.uniform(-1, 1, (s, s)).astype(np.float32)
            a, b, g = T.from_numpy(an), T.from_numpy(bn), T.from_numpy(an.T.copy())
            for mode in modes:
                if mode == 'fp32' and s > 4096:
                    continue                        # 1.1 TFLOP at ~35 TFLOP/s: not worth 100 ms x iterations
                ops.set_matmul_mode(mode)

                def fwd():
                    with light.no_grad():
                        ops._gemm(a, b)

                def fwd_bwd():
                    with light.no_grad():
                        ops._gemm(a, b)
                        ops._gemm(g, ops._swap_last(b))
                        ops._gemm(ops._swap_last(a), g)
                for name, fn, fl in (('fwd', fwd, 2.0), ('fwd+bwd', fwd_bwd, 6.0)):
                    ms = _time(rt, fn, iters=5 if mode != 'fp32' else 3, warmup=2)
                    tf = fl * s ** 3 / ms / 1e9
                    rec = {'op': 'matmul_' + name, 'mode': mode, 'M': s, 'N': s, 'K': s, 'ms': round(ms, 4),
                           'tflops': round(tf, 1), 'frac_of_nominal': round(tf / NOMINAL_TF[mode], 3)}
                    if measured[mode]:
                        rec['frac_of_measured'] = round(tf / measured[mode], 3)
                        rec['measured_peak'] = ('cuBLAS bf16 burst' if mode == 'bf16' else 'half of cuBLAS bf16 burst') + \
                            ' %.1f TFLOP/s' % measured[mode]
                    out.append(rec)
            if cpu and s == sizes[0]:
                dt = _cpu_time(lambda: an @ bn, reps=1)
                out.append({'op': 'matmul_fwd', 'impl': 'cpu-openblas (reference CpuTensor arithmetic)', 'M': s, 'N': s,
                            'K': s, 'ms': round(dt * 1e3, 1), 'tflops': round(2.0 * s ** 3 / dt / 1e12, 3),
                            'cores': os.cpu_count()})
            del a, b, g
    finally:
        ops.set_matmul_mode(prev)
    return out


def mnist_mlp(rt, light, T, cpu):
    import lightgrad_b200.nn as nn
    from examples import mnist as mn
    from lightgrad_b200.autograd.cuda.graph import StepGraph
    out = {}
    with nn.use_tensor(T):
        np.random.seed(0)
        model = mn.NN()
    opt = light.optim.SGD(model.parameters(), lr=1e-4)
    xb, yb = mn.synthetic_batch(64, seed=0)
    xd, yd = T.from_numpy(xb, requires_grad=False), T.from_numpy(yb, requires_grad=False)
    one_hot = T.zeros((64, 10), requires_grad=False)
    one_hot[range(64), yd] = 1

    def step():
        loss = light.loss.mse(model(xd), one_hot)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss
    for _ in range(5):
        step()
    rt.synchronize()
    n0 = rt.launch_count()
    t0 = time.perf_counter()
    for _ in range(100):
        step()
    rt.synchronize()
    dt = (time.perf_counter() - t0) / 100
    out['eager'] = {'ms_per_step': round(dt * 1e3, 4), 'samples_per_s': round(64 / dt, 1),
                    'launches_per_step': (rt.launch_count() - n0) / 100.0}
    sg = StepGraph(step, warmup=0)
    for _ in range(5):
        sg.replay()
    ms = _span(rt, sg.replay, 500)
    out['graph_replay'] = {'ms_per_step': round(ms, 4), 'samples_per_s': round(64 / (ms / 1e3), 1),
                           'kernels_per_step': sg.n_kernels, 'bound': 'launch latency (39 MFLOP, 0.8 MB per step)'}
    del sg
    if cpu:
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
        out['cpu_oracle'] = {'ms_per_step': round(dt * 1e3, 4), 'samples_per_s': round(64 / dt, 1),
                             'cores': os.cpu_count(), 'kind': 'port'}
    return out


def run(rt, light, T, ops, peaks, cpu=True, modes=('fp32', 'tf32')):
    t0 = time.perf_counter()
    res = {'config1_mnist_mlp': mnist_mlp(rt, light, T, cpu),
           'config2_elementwise_reduce': elementwise(rt, light, T, ops, peaks, cpu),
           'config3_matmul': matmul(rt, light, T, ops, peaks, cpu, modes),
           'peaks': {'hbm_gbs_measured': peaks['hbm_gbs'], 'bf16_tflops_burst_measured': peaks['bf16_tflops'],
                     'nominal_tflops': NOMINAL_TF, 'targets': 'elementwise/reduce >= 6000 GB/s (75% of 8 TB/s); '
                                                              'matmul >= 70% of nominal at 4096^3'}}
    res['seconds'] = round(time.perf_counter() - t0, 1)
    return res
