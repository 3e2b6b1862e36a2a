#!/usr/bin/env python
"""bench.py -- BERT-base masked-LM training throughput on lightgrad_b200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU tensor path (oracle port)

A "step" is one pass of the hot path over one batch of synthetic tokens: forward, fused cross
entropy, backward (topological walk), [N>1: gradient exchange over NVLink], Adam update.
Workloads (BASELINE.json configs[3] / configs[4]):
    N = 1   BERT-base (12L, d=768, seq 128), batch 32, random-init weights, synthetic tokens
    N > 1   same model, 32 samples PER GPU (global batch 32 N, rows sharded over the ranks): weak scaling
            against the N = 1 line at every N; at N = 8 this is exactly configs[4] (global batch 256).
            --global-batch 256 runs configs[4] at N = 2 / 4 as written; the default N = 2 / 4 lines
            carry that measurement too, in `config5_global_batch_256`.
One JSON line is printed by rank 0:
    value      samples/s, whole job, inputs resident in HBM before the timed region (CUDA events)
    e2e        the same step driven through the public API with the step's token ids / labels copied
               from pinned host memory and the loss read back to the host EVERY step
    roofline   the dominant kernel (matmul): algorithmic flops / summed CUDA-event kernel time
    cpu_baseline  the oracle port of the reference CPU tensor on this box's host cores (bounded sample)
    ops        N = 1 only: BASELINE.json configs 1-3 (MNIST MLP steps/s, elementwise / reduction GB/s, matmul
               TFLOP/s per mode) measured after the timed region, each next to the CPU tensor's arithmetic
    modes      N = 1 only: the same BERT step in the other matmul modes (exact fp32, tf32, bf16)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if '--impl' in sys.argv and 'reference' in sys.argv:
    # the CPU arm uses every host core: torchrun exports OMP_NUM_THREADS=1 to its workers, which would make
    # OpenBLAS single-threaded; this has to happen before numpy is imported
    for _k in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[_k] = str(os.cpu_count() or 1)
# stdout carries exactly one JSON line.  Libraries write there behind python's back (NCCL prints its version
# banner to fd 1 whatever NCCL_DEBUG_FILE says), so fd 1 is pointed at stderr for the whole run and the JSON
# line goes to a private duplicate of the original stdout.
os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
sys.stdout.flush()
_REAL_STDOUT = os.fdopen(os.dup(1), 'w')
os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + '\n')
    _REAL_STDOUT.flush()


import numpy as np  # noqa: E402

SEQ = 128
GEMM_FLOPS_PER_SAMPLE = 85.5e9   # fwd + both backward GEMMs, SURVEY.md 8(a)/BASELINE.md "work per unit"
# ncu dram__bytes_read.sum + dram__bytes_write.sum over the matmul launches of one batch-32 step on 1 GPU, per mode
NCU_GEMM_DRAM_BYTES_PER_STEP = {
    'tf32': (10.84e9, 'profiles/r2_gemm_dram_bytes.csv: the 150 matmul launches of one batch-32 step, 9.71 GB read + '
                      '1.13 GB written (r1, before the fused attention kernels: 222 launches, 13.02 GB; results still '
                      'in L2 at kernel end are charged to later kernels)'),
}


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        d['source'] = 'measured (MEASURED_PEAKS.json)'
        return d
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0,
            'source': 'fallback (B200_PROFILING.md)'}


class ClockSampler(object):
    """Samples SM clock / throttle reasons with nvidia-smi while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        clocks, mx, reasons, power = [], None, set(), []
        for r in self.rows:
            try:
                clocks.append(float(r[0]))
                mx = float(r[1])
                power.append(float(r[2]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                continue
        clocks.sort()
        return {'sm_mhz': clocks[len(clocks) // 2] if clocks else None, 'sm_max_mhz': mx,
                'power_w_max': max(power) if power else None, 'samples': len(clocks), 'reasons': sorted(reasons)}


def build_bert(T, cfg, seed=0):
    import lightgrad_b200.nn as nn
    from examples import bert
    with nn.use_tensor(T):
        np.random.seed(seed)
        return bert.BertForMaskedLM(**cfg)


def make_step(model, opt, dp, light):
    from examples import bert  # noqa: F401

    def step(ids, labels):
        logits = model(ids)
        loss = light.loss.cross_entropy(logits.reshape(-1, model.vocab_size), labels)
        opt.zero_grad()
        if dp is not None:
            # backward with, per bucket of parameters whose gradients are final, the gradient exchange fused with the
            # optimizer update (N > 1) or the optimizer update alone (N = 1) on the collective stream, beside backward
            dp.backward_and_step(loss)
        else:
            loss.backward()
            opt.step()
        return loss
    return step


def cpu_reference_leg(steps, warmup, sample_batch, budget_s=150.0):
    """The reference's CPU tensor path (oracle port, numpy + OpenBLAS on every host core) on a bounded
    sample of the workload: the same BERT-base model, sequence length and optimizer, batch `sample_batch`."""
    import lightgrad_b200 as light
    from oracle import CpuTensor
    from examples import bert
    model = build_bert(CpuTensor, bert.BERT_BASE)
    opt = light.optim.Adam(model.parameters(), lr=1e-4)
    step = make_step(model, opt, None, light)
    ids, labels = bert.synthetic_batch(sample_batch, SEQ, bert.BERT_BASE['vocab_size'])
    ids_t, lab_t = CpuTensor.from_numpy(ids, requires_grad=False), CpuTensor.from_numpy(labels, requires_grad=False)
    t_start = time.perf_counter()
    done_w = 0
    for _ in range(warmup):
        step(ids_t, lab_t)
        done_w += 1
        if time.perf_counter() - t_start > budget_s * 0.4:
            break
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        loss = step(ids_t, lab_t)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    per = float(np.mean(times))
    return {'value': sample_batch / per, 'unit': 'samples/s', 'cores': os.cpu_count(), 'kind': 'port',
            'sample': 'BERT-base seq %d, batch %d per step (full model + Adam), %d warm-up + %d timed steps, '
                      'numpy/OpenBLAS oracle port of the reference CpuTensor (+sum/dot/getitem backward patches, '
                      'topological walk)' % (SEQ, sample_batch, done_w, len(times)),
            'ms_per_step': per * 1e3, 'steps_timed': len(times), 'loss': float(loss.item())}


# matmul mode of the headline line when --mode auto: the fastest tensor-core mode whose end-to-end gradients are
# held to the north star's 5e-3 by the -m gpu parity tests (tests/test_full_size.py)
DEFAULT_TC_MODE = os.environ.get('LG_BENCH_MODE', 'tf32')
PER_GPU_BATCH = 32


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--mode', default='auto', help='fp32 | tf32 | bf16 | auto')
    ap.add_argument('--per-gpu-batch', type=int, default=0, help='samples per GPU (default 32): weak scaling')
    ap.add_argument('--global-batch', '--batch', dest='global_batch', type=int, default=0,
                    help='fix the GLOBAL batch instead (e.g. 256 = BASELINE configs[4] at any N): strong scaling')
    ap.add_argument('--layers', type=int, default=0, help='debug only: fewer layers (marks the line invalid)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-ops', action='store_true', help='skip the configs 1-3 suite and the other-mode lines')
    ap.add_argument('--eager', action='store_true', help='dispatch every op from python every step (no CUDA graph replay)')
    ap.add_argument('--cpu-sample-batch', type=int, default=0,
                    help='oracle sample batch (default 8 for --impl reference, 4 for the cpu_baseline leg)')
    return ap.parse_args()


def workload_config(args):
    per_gpu = args.per_gpu_batch or PER_GPU_BATCH
    global_batch = args.global_batch or per_gpu * args.gpus
    assert global_batch % args.gpus == 0, "global batch %d does not divide over %d GPUs" % (global_batch, args.gpus)
    weak = not args.global_batch
    config = {'workload': 'examples/bert.py BERT-base (12L, d=768, heads 12, vocab 30522) masked-LM training step, '
                          'seq %d, global batch %d (%d per GPU), Adam(lr=1e-4), random-init weights, synthetic tokens'
                          % (SEQ, global_batch, global_batch // args.gpus),
              'global_batch': global_batch, 'per_gpu_batch': global_batch // args.gpus, 'seq_len': SEQ,
              'parallelism': 'dp%d' % args.gpus,
              'baseline_config': ('configs[3] (1 GPU, batch 32)' if args.gpus == 1 and global_batch == 32 else
                                  'configs[4] (global batch 256)' if global_batch == 256 else
                                  'configs[3] per-GPU workload replicated over %d GPUs' % args.gpus),
              'l2_policy': 'per-step working set (531.8 MB parameters + activations) exceeds the 126 MB L2'}
    return config, global_batch, ('weak' if weak else 'strong')


def reference_arm(args, config, scaling):
    """The reference's CPU implementation of the path (oracle port: numpy + OpenBLAS on every host core) on this
    arm's workload; each step is a bounded sample of it (the full model and optimizer, a smaller batch)."""
    sample = args.cpu_sample_batch or 8
    res = cpu_reference_leg(args.steps, args.warmup, sample, budget_s=170.0)
    config = dict(config)
    config['reference_sample'] = ('each timed step runs batch %d of the workload above (full model, seq %d, Adam); '
                                  'samples/s = %d / step time' % (sample, SEQ, sample))
    config['sample_batch_per_step'] = sample
    config['same_config'] = False
    config['blas_threads'] = os.environ.get('OMP_NUM_THREADS')
    emit({'impl': 'reference', 'metric': 'bert_train_samples_per_s', 'value': res['value'], 'unit': 'samples/s',
          'n_gpus': args.gpus, 'steps': res['steps_timed'], 'warmup': args.warmup,
          'ms_per_step': res['ms_per_step'], 'higher_is_better': True, 'scaling': scaling, 'vs_baseline': None,
          'dtype': 'f32', 'data': 'synthetic', 'config': config,
          'cpu_baseline': {k: res[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
          'e2e': {'value': res['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
          'gpu_launches': 0})


def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    W = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    config, global_batch, scaling = workload_config(args)

    if args.impl == 'reference':
        if rank == 0:
            reference_arm(args, config, scaling)
        return

    # ------------------------------------------------------------------------------------ our arm
    import lightgrad_b200 as light
    from lightgrad_b200 import CudaTensor, parallel
    from lightgrad_b200.autograd.cuda import runtime as rt, ops
    from lightgrad_b200.autograd.cuda.graph import StepGraph
    from examples import bert
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus
    rt.ensure_device(int(os.environ.get('LOCAL_RANK', '0')))
    import ctypes
    probe = rt.GemmDesc(4096, 768, 768, 1, 1, 0, 0, 768, 1, 0, 0, 1, 768, 0, 0, 768, 1)
    mode = args.mode
    if mode == 'auto':
        tc_code = {'tf32': rt.GEMM_TF32_TC, 'bf16': rt.GEMM_BF16_TC}.get(DEFAULT_TC_MODE, rt.GEMM_TF32_TC)
        mode = DEFAULT_TC_MODE if rt.api.gemm_tc_supported(tc_code, rt.F32, ctypes.byref(probe)) else 'fp32'
    ops.set_matmul_mode(mode)
    cfg = dict(bert.BERT_BASE)
    if args.layers:
        cfg['num_hidden_layers'] = args.layers
        config['INVALID'] = 'debug run with %d layers' % args.layers
    model = build_bert(CudaTensor, cfg)
    opt = light.optim.Adam(model.parameters(), lr=1e-4)
    comm = parallel.default_comm() if world > 1 else parallel.LocalComm()
    # one GPU: loss.backward(); optimizer.step().  (LG_BENCH_OVERLAP_STEP=1 runs the optimizer per bucket beside backward
    # instead, DataParallel exchange='local': measured slower, 8.97 vs 8.53 ms, see parallel.py)
    if world > 1:
        dp = parallel.DataParallel(model, opt, comm=comm)
    elif os.environ.get('LG_BENCH_OVERLAP_STEP'):
        dp = parallel.DataParallel(model, opt, comm=comm, exchange='local')
    else:
        dp = None
    step = make_step(model, opt, dp, light)
    light.Gradients.retain_intermediate = False
    graphs = []                                  # every captured step, destroyed in order at teardown

    def barrier():
        rt.synchronize()
        comm.barrier()

    def shard_inputs(batch):
        ids_g, labels_g = bert.synthetic_batch(batch, SEQ, cfg['vocab_size'])
        lo, hi = parallel.shard_rows(batch, rank, world)
        ids_np, labels_np = ids_g[lo:hi], labels_g[lo * SEQ:hi * SEQ]
        return (ids_np, labels_np, CudaTensor.from_numpy(ids_np, requires_grad=False),
                CudaTensor.from_numpy(labels_np, requires_grad=False))

    def timed(run, k):
        """ms per call of run(): k calls between two CUDA events on the compute stream, ranks released together,
        max over ranks."""
        run()
        barrier()
        t0 = rt.Event().record()
        out = None
        for _ in range(k):
            out = run()
        t1 = rt.Event().record()
        t1.synchronize()
        barrier()
        return comm.max_float(t0.elapsed_ms(t1)) / k, out

    def captured(step_fn, ids, labels, warm):
        sg = StepGraph(lambda: step_fn(ids, labels), warmup=0)
        graphs.append(sg)
        for _ in range(warm):
            sg.replay()
        return sg

    ids_np, labels_np, ids_d, lab_d = shard_inputs(global_batch)
    local = ids_np.shape[0]

    # ---- eager reference point: python dispatch of every op, every step
    for _ in range(2):
        loss = step(ids_d, lab_d)
    eager_ms, _ = timed(lambda: step(ids_d, lab_d), 3)

    # ---- the step is static-shape: record it once into a CUDA graph, replay it without python dispatch
    if args.eager:
        run_step = lambda: step(ids_d, lab_d)  # noqa: E731
        for _ in range(W):
            run_step()
    else:
        sg = captured(step, ids_d, lab_d, W)
        run_step = sg.replay

    # ---- phase A: inputs resident in HBM
    barrier()
    sampler = ClockSampler(int(os.environ.get('LOCAL_RANK', '0')))
    if rank == 0:
        sampler.start()
    n0 = rt.launch_count()
    e0 = rt.Event().record()
    for _ in range(args.steps):
        loss = run_step()
    e1 = rt.Event().record()
    e1.synchronize()
    n1 = rt.launch_count()
    barrier()
    ms_a = comm.max_float(e0.elapsed_ms(e1)) / args.steps
    final_loss = float(loss.item())

    # ---- phase B: end to end -- pinned host inputs copied in and the loss read back every step
    pin_ids, pin_lab = rt.PinnedArray(ids_np.shape, np.int32), rt.PinnedArray(labels_np.shape, np.int32)
    pin_ids.array[...] = ids_np
    pin_lab.array[...] = labels_np
    k_b = max(3, min(args.steps, 10))
    barrier()
    e2 = rt.Event().record()
    for _ in range(k_b):
        rt.api.memcpy_h2d(ids_d.ptr, pin_ids.ptr, ids_np.nbytes)
        rt.api.memcpy_h2d(lab_d.ptr, pin_lab.ptr, labels_np.nbytes)
        loss_host = run_step().item()
    e3 = rt.Event().record()
    e3.synchronize()
    barrier()
    ms_b = comm.max_float(e2.elapsed_ms(e3)) / k_b
    clocks = sampler.stop() if rank == 0 else None

    # ---- phase C: device time of the dominant kernel (matmul) inside the step.
    # Event pairs around every launch serialise neighbouring kernels (~9 us per small GEMM), so with graph
    # replay the matmul time is measured by ablation instead: the same step is captured a second time with
    # lg_gemm launching nothing (lg_prof_gemm(2); values are garbage, the other kernels and their order are
    # identical) and both graphs are timed with CUDA events on the compute stream; the difference is what
    # the matmul launches cost inside the replayed step.  (--eager: event pairs around every launch, the
    # step queued behind a stream delay so the events see no host dispatch latency.)
    k_c = 5
    rt.gemm_profile_read()
    if args.eager:
        rt.gemm_profile(1)
        rt.synchronize()
        rt.api.stream_delay_us(int(3 * eager_ms * 1000) + 20000)
        e4 = rt.Event().record()
        step(ids_d, lab_d)
        e5 = rt.Event().record()
        e5.synchronize()
        gemm_ms, gemm_launches, gemm_flops = rt.gemm_profile_read()
        step_ms_c = e4.elapsed_ms(e5)
        roof_how = ('CUDA events on the compute stream around every lg_gemm launch of one extra step queued behind '
                    'a stream delay (no host dispatch latency inside the events)')
    else:
        step_ms_c, _ = timed(sg.replay, k_c)
        rt.gemm_profile(2)
        sg_nogemm = captured(step, ids_d, lab_d, 0)
        _, gemm_launches, gemm_flops = rt.gemm_profile_read()
        rt.gemm_profile(0)
        nogemm_ms, _ = timed(sg_nogemm.replay, k_c)
        gemm_ms = max(step_ms_c - nogemm_ms, 1e-6)
        roof_how = ('ablation inside the replayed CUDA graph: %d replays of the step (%.3f ms) minus %d replays of the '
                    'same captured step with the matmul launches removed (%.3f ms), both timed with CUDA events on '
                    'the compute stream right after the timed region' % (k_c, step_ms_c, k_c, nogemm_ms))
        sg.replay()                                # leave real values behind the garbage-valued ablation replays
    rt.gemm_profile(False)

    # ---- multi-GPU only: the gradient exchange on its own and what of it stays exposed
    comm_info, config5 = None, None
    if world > 1 and dp is not None:
        comm_info = dp.exchange_report(rt, comm, barrier) if hasattr(dp, 'exchange_report') else {}
        comm_info['exchange'] = dp.exchange
        if dp.exchange_note:
            comm_info['exchange_note'] = dp.exchange_note
        # the same local step without any exchange (captured separately): step time minus this = exposed communication
        local_step = make_step(model, opt, None, light)
        if args.eager:
            local_ms, _ = timed(lambda: local_step(ids_d, lab_d), 3)
        else:
            local_ms, _ = timed(captured(local_step, ids_d, lab_d, 1).replay, k_c)
        comm_info.update({'step_ms_without_exchange': round(local_ms, 3),
                          'exposed_comm_ms': round(max(ms_a - local_ms, 0.0), 3),
                          'like_for_like_efficiency': round(local_ms / ms_a, 4),
                          'like_for_like_note': 'this rank count\'s step time without any exchange (same local batch, '
                                                'same binary, same box) divided by the step time with it'})
        if not args.global_batch and global_batch != 256 and not args.eager and 256 % world == 0:
            # BASELINE configs[4] exactly as written, at this N: global batch 256
            _i5, _l5, ids5, lab5 = shard_inputs(256)
            step(ids5, lab5)                      # eager once: shape-dependent constants are built outside the capture
            ms5, _ = timed(captured(step, ids5, lab5, 3).replay, 10)
            config5 = {'global_batch': 256, 'per_gpu_batch': 256 // world, 'ms_per_step': round(ms5, 3),
                       'value': round(256 / (ms5 / 1e3), 2), 'unit': 'samples/s', 'scaling': 'strong', 'steps': 10}
            del ids5, lab5

    # ---- N = 1: the same step in the other matmul modes, and BASELINE configs 1-3
    modes_info, ops_info = None, None
    if args.gpus == 1 and not args.no_ops and not args.eager and not args.layers:
        modes_info = {}
        for other in ('fp32', 'tf32', 'bf16'):
            if other == mode:
                continue
            code = {'tf32': rt.GEMM_TF32_TC, 'bf16': rt.GEMM_BF16_TC}.get(other)
            if code is not None and not rt.api.gemm_tc_supported(code, rt.F32, ctypes.byref(probe)):
                modes_info[other] = {'unavailable': 'lg_gemm_tc_supported says no for this mode'}
                continue
            try:
                ops.set_matmul_mode(other)
                sgo = captured(step, ids_d, lab_d, 3)
                ms_o, loss_o = timed(sgo.replay, 5 if other != 'fp32' else 3)
                modes_info[other] = {'ms_per_step': round(ms_o, 3), 'value': round(global_batch / (ms_o / 1e3), 2),
                                     'unit': 'samples/s', 'loss': round(float(loss_o.item()), 5)}
            except Exception as exc:                       # a mode that is not built must not take the line down
                modes_info[other] = {'unavailable': '%s: %s' % (type(exc).__name__, str(exc)[:200])}
            finally:
                ops.set_matmul_mode(mode)
        # release the step graphs' private pools before the 2^28-element sweeps
        for g in graphs:
            g.destroy()
        del graphs[:]
        sys.path.insert(0, os.path.join(ROOT, 'benchmarks'))
        import ops_suite
        peaks_ = load_peaks()
        tc_modes = ['fp32', 'tf32'] + (['bf16'] if 'unavailable' not in (modes_info.get('bf16') or {}) or mode == 'bf16' else [])
        ops_info = ops_suite.run(rt, light, CudaTensor, ops, peaks_, cpu=not args.no_cpu_baseline, modes=tc_modes)

    if rank == 0:
        peaks = load_peaks()
        value = global_batch / (ms_a / 1e3)
        e2e = global_batch / (ms_b / 1e3)
        achieved_tf = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        # a sub-second timed region at full clocks: the burst figures are the honest denominators
        burst, sustained = peaks['bf16_tflops'], peaks.get('bf16_tflops_sustained', peaks['bf16_tflops'])
        if mode == 'bf16':
            peak_tf, peak_name, nominal = burst, 'measured cuBLAS bf16 burst', 2250.0
            alt = {'frac_of_sustained_bf16': round(achieved_tf / sustained, 4)}
        elif mode == 'tf32':
            peak_tf, nominal = burst / 2.0, 1125.0
            peak_name = 'half of the measured cuBLAS bf16 burst (tf32 runs at half the bf16 rate; no tf32 figure in MEASURED_PEAKS.json)'
            alt = {'frac_of_half_sustained_bf16': round(achieved_tf / (sustained / 2.0), 4)}
        else:
            peak_tf = nominal = 148 * 128 * 2 * 1.965e9 / 1e12
            peak_name, alt = 'nominal FP32 FMA pipe (148 SM x 128 lanes x 2 x 1.965 GHz): exact-fp32 SIMT mode', {}
        traffic = NCU_GEMM_DRAM_BYTES_PER_STEP.get(mode) if (args.gpus == 1 and local == 32 and not args.layers) else None
        roofline = {'bound': 'tensor', 'kernel': 'lg_gemm (%s)' % mode, 'achieved': round(achieved_tf, 2),
                    'peak': round(peak_tf, 1), 'unit': 'TFLOP/s', 'frac': round(achieved_tf / peak_tf, 4),
                    'frac_nominal': round(achieved_tf / nominal, 4), 'nominal_peak': nominal,
                    'peak_source': peak_name + '; ' + peaks['source'],
                    # DRAM bytes of ALL matmul launches of one step (like `achieved`, an aggregate over the step's
                    # launches): ncu dram__bytes_read.sum + dram__bytes_write.sum over those launches
                    'traffic': traffic[0] if traffic else None,
                    'traffic_note': traffic[1] if traffic else 'no ncu capture for this mode / batch',
                    'launches_per_step': int(gemm_launches), 'kernel_ms_per_step': round(gemm_ms, 3),
                    'kernel_share_of_step': round(gemm_ms / step_ms_c, 3),
                    'algorithmic_flops_per_step': gemm_flops, 'how': roof_how}
        roofline.update(alt)
        line = {
            'metric': 'bert_train_samples_per_s', 'value': round(value, 2), 'unit': 'samples/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': W, 'ms_per_step': round(ms_a, 3), 'higher_is_better': True,
            'scaling': scaling, 'vs_baseline': None,
            'dtype': {'fp32': 'f32', 'tf32': 'tf32', 'bf16': 'bf16'}[mode], 'data': 'synthetic', 'config': config,
            'loss': round(final_loss, 5),
            'execution': 'eager python dispatch' if args.eager else 'whole step captured once into a CUDA graph, replayed per step',
            'optimizer_step': ('per-bucket updates on the collective stream, overlapped with backward (exchange %r)' % dp.exchange)
            if dp is not None else 'optimizer.step() after backward',
            'eager_ms_per_step': round(eager_ms, 3),
            'e2e': {'value': round(e2e, 2), 'unit': 'samples/s', 'h2d_bytes_per_step': int(ids_np.nbytes + labels_np.nbytes),
                    'd2h_bytes_per_step': 4, 'ms_per_step': round(ms_b, 3), 'steps': k_b, 'last_loss': round(float(loss_host), 5)},
            'gpu_launches': int(n1 - n0), 'gpu_launches_per_step': round((n1 - n0) / args.steps, 1),
            'clocks': clocks, 'roofline': roofline,
            # whole-step model flops utilisation PER GPU (the aggregate divided by the GPUs that produced it)
            'mfu_vs_measured_bf16': round(GEMM_FLOPS_PER_SAMPLE * value / args.gpus / 1e12 / peaks['bf16_tflops'], 4),
        }
        if comm_info is not None:
            line['comm'] = comm_info
        if config5 is not None:
            line['config5_global_batch_256'] = config5
        if modes_info is not None:
            line['modes'] = modes_info
        if ops_info is not None:
            line['ops'] = ops_info
        if args.gpus == 1 and not args.no_cpu_baseline:
            res = cpu_reference_leg(2, 1, args.cpu_sample_batch or 4, budget_s=90.0)
            line['cpu_baseline'] = {k: res[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
        emit(line)

    # ---- orderly teardown (normal interpreter exit afterwards): captured steps first, then the communicator
    barrier()
    for g in graphs:
        g.destroy()
    del graphs[:]
    if dp is not None:
        dp.close()
    comm.close()


if __name__ == '__main__':
    main()
