"""SURVEY.md 8(d) config 5 parity rule on real GPUs: N-rank averaged gradients == 1-rank gradients on the same
global batch.  (The parameters after the first Adam step are NOT compared: that step is lr * sign(g) for every
element, so an element whose gradient is ~0 can legitimately move by +lr on one side and -lr on the other.)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 \
        benchmarks/dp_parity.py [--mode fp32|tf32]

Every rank runs the data-parallel step (bucketed all-reduce overlapped with backward, weight gradients on the side
stream) on its shard; rank 0 then repeats the step alone on the whole batch with an identically initialised model and
compares.  Prints one JSON line; exit status 1 on a mismatch.
"""
import argparse
import json
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lightgrad_b200 as light                                         # noqa: E402
import lightgrad_b200.nn as nn                                         # noqa: E402
from lightgrad_b200 import CudaTensor, parallel                        # noqa: E402
from lightgrad_b200.autograd.cuda import ops, runtime as rt            # noqa: E402
from examples import bert                                              # noqa: E402

CFG = dict(hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, vocab_size=1000,
           max_position_embeddings=64, type_vocab_size=2)


def build():
    with nn.use_tensor(CudaTensor):
        np.random.seed(0)
        model = bert.BertForMaskedLM(**CFG)
    return model, light.optim.Adam(model.parameters(), lr=1e-3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mode', default='fp32')
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--seq', type=int, default=32)
    a = ap.parse_args()
    rt.ensure_device(int(os.environ.get('LOCAL_RANK', '0')))
    ops.set_matmul_mode(a.mode)
    comm = parallel.default_comm()
    rank, world = comm.rank, comm.world
    ids, labels = bert.synthetic_batch(a.batch, a.seq, CFG['vocab_size'])

    model, opt = build()
    dp = parallel.DataParallel(model, opt, comm=comm)
    lo, hi = dp.shard(a.batch)
    x = CudaTensor.from_numpy(ids[lo:hi], requires_grad=False)
    y = CudaTensor.from_numpy(labels[lo * a.seq:hi * a.seq], requires_grad=False)
    loss = light.loss.cross_entropy(model(x).reshape(-1, CFG['vocab_size']), y)
    opt.zero_grad()
    dp.backward(loss)
    grads = [p.grad.numpy().copy() for p in model.parameters()]
    opt.step()
    rt.synchronize()
    comm.barrier()
    ok = True
    if rank == 0:
        ref_model, ref_opt = build()
        xf = CudaTensor.from_numpy(ids, requires_grad=False)
        yf = CudaTensor.from_numpy(labels, requires_grad=False)
        ref_loss = light.loss.cross_entropy(ref_model(xf).reshape(-1, CFG['vocab_size']), yf)
        ref_opt.zero_grad()
        ref_loss.backward()
        ref_grads = [p.grad.numpy().copy() for p in ref_model.parameters()]
        gmax = max(float(np.abs(g).max()) for g in ref_grads)
        g_err = max(float(np.abs(g - r).max()) for g, r in zip(grads, ref_grads)) / gmax
        tol = 1e-5 if a.mode == 'fp32' else 5e-3
        ok = g_err <= tol
        print(json.dumps({'world': world, 'mode': a.mode, 'global_batch': a.batch, 'grad_rel_err': g_err,
                          'tol': tol, 'ok': bool(ok)}), flush=True)
    comm.barrier()
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == '__main__':
    main()
