"""Timeline of the NVLink-multicast gradient exchange inside one replayed BERT-base training step.

    LG_MC_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29530 benchmarks/exchange_trace.py [--batch 32] [--out profiles/r2_exchange_trace_nN.json]

The step is captured into a CUDA graph with time stamps (GPU global timer) on the compute stream at the start of
forward, the start of backward and the end of the step, and inside every exchange kernel (entered / all ranks met /
finished).  After a few replays the last replay's records are printed relative to the start of that step: they show
where in backward each bucket's exchange ran, and how much of it lies after the end of backward."""
import argparse
import ctypes as C
import json
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lightgrad_b200 as light                                         # noqa: E402
import lightgrad_b200.nn as nn                                         # noqa: E402
from lightgrad_b200 import CudaTensor, parallel                        # noqa: E402
from lightgrad_b200.autograd.cuda import ops, runtime as rt            # noqa: E402
from lightgrad_b200.autograd.cuda.graph import StepGraph               # noqa: E402
from examples import bert                                              # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32, help='samples per GPU')
    ap.add_argument('--mode', default='tf32')
    ap.add_argument('--out', default='')
    a = ap.parse_args()
    rt.ensure_device(int(os.environ.get('LOCAL_RANK', '0')))
    ops.set_matmul_mode(a.mode)
    comm = parallel.default_comm()
    with nn.use_tensor(CudaTensor):
        np.random.seed(0)
        model = bert.BertForMaskedLM(**bert.BERT_BASE)
    opt = light.optim.Adam(model.parameters(), lr=1e-4)
    dp = parallel.DataParallel(model, opt, comm=comm)
    light.Gradients.retain_intermediate = False
    ids, labels = bert.synthetic_batch(a.batch, 128, 30522, seed=1 + comm.rank)
    x, y = CudaTensor.from_numpy(ids, requires_grad=False), CudaTensor.from_numpy(labels, requires_grad=False)
    api = rt.api

    def step():
        api.mc_trace_mark()                                  # record 0: forward starts
        loss = light.loss.cross_entropy(model(x).reshape(-1, 30522), y)
        opt.zero_grad()
        api.mc_trace_mark()                                  # record 1: backward starts
        dp.backward_and_step(loss)
        api.mc_trace_mark()                                  # last record: the compute stream has joined the exchange
        return loss
    for _ in range(2):
        step()
    n = C.c_int(0)
    buf = (C.c_uint64 * (4 * 4096))()
    api.mc_trace_read(buf, 4096, C.byref(n), 1)
    sg = StepGraph(step, warmup=0)
    for _ in range(5):
        sg.replay()
    rt.synchronize()
    comm.barrier()
    api.mc_trace_read(buf, 4096, C.byref(n), 0)
    rec = np.frombuffer(buf, dtype=np.uint64, count=4 * n.value).reshape(-1, 4).astype(np.int64)
    if comm.rank == 0 and len(rec):
        t0 = rec[0, 0]
        marks = [r for r in rec if r[1] == 0 and r[2] == 0]
        buckets = [r for r in rec if r[2] != 0]
        end_step = marks[-1][0]
        out = {'exchange': dp.exchange, 'world': comm.world, 'per_gpu_batch': a.batch, 'mode': a.mode,
               'forward_starts_us': 0.0, 'backward_starts_us': round((marks[1][0] - t0) / 1e3, 1),
               # the compute stream's last backward kernel has run (stamp queued right after loss.backward())
               'backward_compute_ends_us': round((marks[2][0] - t0) / 1e3, 1) if len(marks) >= 4 else None,
               'step_ends_us': round((end_step - t0) / 1e3, 1),
               'buckets': [{'mbytes': round(r[3] / 1e6, 1), 'entered_us': round((r[0] - t0) / 1e3, 1),
                            'all_ranks_met_us': round((r[1] - t0) / 1e3, 1), 'finished_us': round((r[2] - t0) / 1e3, 1),
                            'busy_us': round((r[2] - r[1]) / 1e3, 1)} for r in buckets]}
        if buckets:
            out['exchange_busy_total_us'] = round(sum(b['busy_us'] for b in out['buckets']), 1)
            out['last_bucket_finishes_us'] = max(b['finished_us'] for b in out['buckets'])
        print(json.dumps(out, indent=1))
        if a.out:
            json.dump(out, open(a.out, 'w'), indent=1)
    sg.destroy()
    dp.close()
    comm.close()


if __name__ == '__main__':
    main()
