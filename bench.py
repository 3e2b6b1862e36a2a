#!/usr/bin/env python
"""bench.py -- BERT-base masked-LM training throughput on lightgrad_b200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU tensor path (oracle port)

A "step" is one pass of the hot path over one batch of synthetic tokens: forward, fused cross
entropy, backward (topological walk), [N>1: one NCCL averaging all-reduce of the flat gradient
arena], Adam update.  Workloads (BASELINE.json configs[3] / configs[4]):
    N = 1   BERT-base (12L, d=768, seq 128), batch 32, random-init weights, synthetic tokens
    N > 1   same model, global batch 256 sharded by rows over the N ranks
One JSON line is printed by rank 0:
    value      samples/s, whole job, inputs resident in HBM before the timed region (CUDA events)
    e2e        the same step driven through the public API with the step's token ids / labels copied
               from pinned host memory and the loss read back to the host EVERY step
    roofline   the dominant kernel (matmul): algorithmic flops / summed CUDA-event kernel time
    cpu_baseline  the oracle port of the reference CPU tensor on this box's host cores (bounded sample)
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
NCU_GEMM_DRAM_BYTES_PER_STEP = 13.02e9   # profiles/r1_gemm_dram_bytes.csv (batch 32, tf32 mode, 1 GPU)


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
            # backward with the bucketed gradient all-reduce overlapped (+ LG_DP_PIPELINED_STEP=1: the optimizer
            # pipelined behind each bucket's exchange), then the optimizer step
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--mode', default=os.environ.get('LG_BENCH_MODE', 'auto'), help='fp32 | tf32 | bf16 | auto')
    ap.add_argument('--batch', type=int, default=0, help='override the global batch')
    ap.add_argument('--layers', type=int, default=0, help='debug only: fewer layers (marks the line invalid)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--eager', action='store_true', help='dispatch every op from python every step (no CUDA graph replay)')
    ap.add_argument('--cpu-sample-batch', type=int, default=0, help='oracle sample batch (default 8 for --impl reference, 4 for the cpu_baseline leg)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    W = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    global_batch = args.batch or (32 if args.gpus == 1 else 256)
    config = {'workload': 'examples/bert.py BERT-base (12L, d=768, heads 12, vocab 30522) masked-LM training step, '
                          'seq %d, global batch %d, Adam(lr=1e-4), random-init weights, synthetic tokens' % (SEQ, global_batch),
              'global_batch': global_batch, 'seq_len': SEQ, 'parallelism': 'dp%d' % args.gpus,
              'l2_policy': 'per-step working set (531.8 MB parameters + activations) exceeds the 126 MB L2'}

    if args.impl == 'reference':
        if rank != 0:
            return
        res = cpu_reference_leg(args.steps, args.warmup, args.cpu_sample_batch or 8, budget_s=170.0)
        line = {'impl': 'reference', 'metric': 'bert_train_samples_per_s', 'value': res['value'], 'unit': 'samples/s',
                'n_gpus': args.gpus, 'steps': res['steps_timed'], 'warmup': args.warmup,
                'ms_per_step': res['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'f32', 'data': 'synthetic', 'config': config,
                'cpu_baseline': {k: res[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
                'e2e': {'value': res['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'gpu_launches': 0}
        emit(line)
        return

    # ------------------------------------------------------------------------------------ our arm
    import lightgrad_b200 as light
    from lightgrad_b200 import CudaTensor, parallel
    from lightgrad_b200.autograd.cuda import runtime as rt, ops
    from examples import bert
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus
    rt.ensure_device(int(os.environ.get('LOCAL_RANK', '0')))
    mode = args.mode
    if mode == 'auto':
        probe = rt.GemmDesc(4096, 768, 768, 1, 1, 0, 0, 768, 1, 0, 0, 1, 768, 0, 0, 768, 1)
        import ctypes
        mode = 'tf32' if rt.api.gemm_tc_supported(rt.GEMM_TF32_TC, rt.F32, ctypes.byref(probe)) else 'fp32'
    ops.set_matmul_mode(mode)
    cfg = dict(bert.BERT_BASE)
    if args.layers:
        cfg['num_hidden_layers'] = args.layers
        config['INVALID'] = 'debug run with %d layers' % args.layers
    model = build_bert(CudaTensor, cfg)
    opt = light.optim.Adam(model.parameters(), lr=1e-4)
    comm = parallel.default_comm() if world > 1 else parallel.LocalComm()
    dp = parallel.DataParallel(model, opt, comm=comm) if world > 1 else None
    step = make_step(model, opt, dp, light)
    light.Gradients.retain_intermediate = False
    ids_g, labels_g = bert.synthetic_batch(global_batch, SEQ, cfg['vocab_size'])
    lo, hi = parallel.shard_rows(global_batch, rank, world)
    local = hi - lo
    ids_np, labels_np = ids_g[lo:hi], labels_g[lo * SEQ:hi * SEQ]
    ids_d = CudaTensor.from_numpy(ids_np, requires_grad=False)
    lab_d = CudaTensor.from_numpy(labels_np, requires_grad=False)

    def barrier():
        rt.synchronize()
        comm.barrier()

    # ---- eager reference point: python dispatch of every op, every step
    for _ in range(2):
        loss = step(ids_d, lab_d)
    barrier()
    ee0 = rt.Event().record()
    for _ in range(3):
        step(ids_d, lab_d)
    ee1 = rt.Event().record()
    ee1.synchronize()
    eager_ms = ee0.elapsed_ms(ee1) / 3

    # ---- the step is static-shape: record it once into a CUDA graph, replay it without python dispatch
    if args.eager:
        run_step = lambda: step(ids_d, lab_d)  # noqa: E731
        for _ in range(W):
            run_step()
    else:
        from lightgrad_b200.autograd.cuda.graph import StepGraph
        sg = StepGraph(lambda: step(ids_d, lab_d), warmup=0)
        run_step = sg.replay
        for _ in range(W):
            run_step()

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
        def timed(replay):
            replay()
            barrier()                       # ranks enter the timed replays together
            t0 = rt.Event().record()
            for _ in range(k_c):
                replay()
            t1 = rt.Event().record()
            t1.synchronize()
            return comm.max_float(t0.elapsed_ms(t1)) / k_c
        step_ms_c = timed(sg.replay)
        rt.gemm_profile(2)
        sg_nogemm = StepGraph(lambda: step(ids_d, lab_d), warmup=0)
        _, gemm_launches, gemm_flops = rt.gemm_profile_read()
        rt.gemm_profile(0)
        nogemm_ms = timed(sg_nogemm.replay)
        gemm_ms = max(step_ms_c - nogemm_ms, 1e-6)
        roof_how = ('ablation inside the replayed CUDA graph: %d replays of the step (%.3f ms) minus %d replays of the '
                    'same captured step with the lg_gemm launches removed (%.3f ms), both timed with CUDA events on '
                    'the compute stream right after the timed region' % (k_c, step_ms_c, k_c, nogemm_ms))
    rt.gemm_profile(False)

    # ---- multi-GPU only: the gradient exchange on its own and what of it stays exposed
    comm_info = None
    if world > 1 and dp is not None and getattr(dp, '_nccl', False):
        a = dp.arena
        nbytes = a.total * 4
        for _ in range(2):
            rt.api.nccl_allreduce_f32(a.grad_buf.ptr, a.total, 1, 0)
        barrier()
        c0 = rt.Event().record()
        for _ in range(5):
            rt.api.nccl_allreduce_f32(a.grad_buf.ptr, a.total, 1, 0)
        c1 = rt.Event().record()
        c1.synchronize()
        ar_ms = comm.max_float(c0.elapsed_ms(c1)) / 5
        # the same local step without any exchange (captured separately): step time minus this = exposed communication
        local_step = make_step(model, opt, None, light)
        if args.eager:
            run_local = lambda: local_step(ids_d, lab_d)  # noqa: E731
        else:
            sg_local = StepGraph(lambda: local_step(ids_d, lab_d), warmup=0)
            run_local = sg_local.replay
        run_local()
        barrier()
        l0 = rt.Event().record()
        for _ in range(k_c if not args.eager else 3):
            run_local()
        l1 = rt.Event().record()
        l1.synchronize()
        local_ms = comm.max_float(l0.elapsed_ms(l1)) / (k_c if not args.eager else 3)
        comm_info = {'allreduce_bytes': int(nbytes), 'allreduce_ms_alone': round(ar_ms, 3),
                     'allreduce_busbw_gbps': round(2.0 * (world - 1) / world * nbytes / (ar_ms / 1e3) / 1e9, 1),
                     'step_ms_without_exchange': round(local_ms, 3),
                     'exposed_comm_ms': round(max(ms_a - local_ms, 0.0), 3),
                     'note': 'bucketed all-reduce (64 MB buckets) overlapped with backward on the communication stream'}
    if rank != 0:
        finish(comm)

    peaks = load_peaks()
    value = global_batch / (ms_a / 1e3)
    e2e = global_batch / (ms_b / 1e3)
    achieved_tf = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    k_c = 1   # gemm_ms / gemm_launches / gemm_flops are per step from here on
    if mode == 'bf16':
        peak_tf, peak_name = peaks.get('bf16_tflops_sustained', peaks['bf16_tflops']), 'measured cuBLAS bf16 (sustained)'
    elif mode == 'tf32':
        peak_tf = peaks.get('bf16_tflops_sustained', peaks['bf16_tflops']) / 2.0
        peak_name = 'half of measured cuBLAS bf16 sustained (tf32 runs at half the bf16 rate; no tf32 figure in MEASURED_PEAKS.json)'
    else:
        peak_tf, peak_name = 148 * 128 * 2 * 1.965e9 / 1e12, 'nominal FP32 FMA pipe (148 SM x 128 lanes x 2 x 1.965 GHz): exact-fp32 SIMT mode'
    line = {
        'metric': 'bert_train_samples_per_s', 'value': round(value, 2), 'unit': 'samples/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': W, 'ms_per_step': round(ms_a, 3), 'higher_is_better': True,
        'scaling': 'weak' if args.gpus == 1 else 'strong', 'vs_baseline': None,
        'dtype': {'fp32': 'f32', 'tf32': 'tf32', 'bf16': 'bf16'}[mode], 'data': 'synthetic', 'config': config,
        'loss': round(final_loss, 5),
        'execution': 'eager python dispatch' if args.eager else 'whole step captured once into a CUDA graph, replayed per step',
        'eager_ms_per_step': round(eager_ms, 3),
        'e2e': {'value': round(e2e, 2), 'unit': 'samples/s', 'h2d_bytes_per_step': int(ids_np.nbytes + labels_np.nbytes),
                'd2h_bytes_per_step': 4, 'ms_per_step': round(ms_b, 3), 'steps': k_b, 'last_loss': round(float(loss_host), 5)},
        'gpu_launches': int(n1 - n0), 'gpu_launches_per_step': round((n1 - n0) / args.steps, 1),
        'clocks': clocks,
        'roofline': {'bound': 'tensor', 'kernel': 'lg_gemm (%s)' % mode, 'achieved': round(achieved_tf, 2),
                     'peak': round(peak_tf, 1), 'unit': 'TFLOP/s', 'frac': round(achieved_tf / peak_tf, 4),
                     'peak_source': peak_name + '; ' + peaks['source'],
                     # DRAM bytes of ALL matmul launches of one step (like `achieved`, an aggregate over the step's
                     # launches): ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r1_gemm_dram_bytes.csv
                     'traffic': NCU_GEMM_DRAM_BYTES_PER_STEP if (mode == 'tf32' and args.gpus == 1 and not args.layers) else None,
                     'traffic_note': 'ncu capture of the 222 matmul launches of one batch-32 step: 11.90 GB read + 1.12 GB '
                                     'written (algorithmic operand + result bytes of those launches: ~16 GB; results '
                                     'still in L2 at kernel end are charged to later kernels)',
                     'launches_per_step': gemm_launches // k_c, 'kernel_ms_per_step': round(gemm_ms / k_c, 3),
                     'kernel_share_of_step': round(gemm_ms / k_c / step_ms_c, 3),
                     'algorithmic_flops_per_step': gemm_flops / k_c,
                     'how': roof_how},
        'mfu_vs_measured_bf16': round(GEMM_FLOPS_PER_SAMPLE * value / 1e12 / peaks['bf16_tflops'], 4),
    }
    if comm_info is not None:
        line['comm'] = comm_info
    line['config']['per_gpu_batch'] = int(local)
    if args.gpus == 1 and not args.no_cpu_baseline:
        res = cpu_reference_leg(2, 1, args.cpu_sample_batch or 4, budget_s=90.0)
        line['cpu_baseline'] = {k: res[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
    emit(line)
    finish(comm)


def finish(comm):
    """Leave without running destructors: a NCCL communicator that was captured into a CUDA graph must not
    be torn down rank by rank (ncclCommDestroy can wait for peers that have already gone)."""
    sys.stdout.flush()
    sys.stderr.flush()
    try:
        comm.barrier()
    except Exception:
        pass
    os._exit(0)


if __name__ == '__main__':
    main()
