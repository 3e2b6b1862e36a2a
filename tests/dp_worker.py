"""Worker of tests/test_parallel_gpu.py: one rank of a 2-GPU data-parallel step (launched with RANK / WORLD_SIZE /
LOCAL_RANK / MASTER_* in the environment).  Writes what rank 0 measured to the JSON file named on the command line
and leaves through a normal interpreter exit (DataParallel.close, no os._exit)."""
import json
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lightgrad_b200 as light                                         # noqa: E402
import lightgrad_b200.nn as nn                                         # noqa: E402
from lightgrad_b200 import CudaTensor, parallel                        # noqa: E402
from lightgrad_b200.autograd.cuda import ops, runtime as rt            # noqa: E402
from examples import bert                                              # noqa: E402

CFG = dict(hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, vocab_size=1000,
           max_position_embeddings=64, type_vocab_size=2)
BATCH, SEQ, STEPS = 16, 32, 3


def build():
    with nn.use_tensor(CudaTensor):
        np.random.seed(0)
        model = bert.BertForMaskedLM(**CFG)
    return model, light.optim.Adam(model.parameters(), lr=1e-3)


def loss_of(model, x, y):
    return light.loss.cross_entropy(model(x).reshape(-1, CFG['vocab_size']), y)


def main():
    out_path, mode = sys.argv[1], sys.argv[2]
    rt.ensure_device(int(os.environ.get('LOCAL_RANK', '0')))
    ops.set_matmul_mode(mode)
    comm = parallel.default_comm()
    rank, world = comm.rank, comm.world
    ids, labels = bert.synthetic_batch(BATCH, SEQ, CFG['vocab_size'])

    model, opt = build()
    dp = parallel.DataParallel(model, opt, comm=comm)
    lo, hi = dp.shard(BATCH)
    x = CudaTensor.from_numpy(ids[lo:hi], requires_grad=False)
    y = CudaTensor.from_numpy(labels[lo * SEQ:hi * SEQ], requires_grad=False)
    # (1) averaged gradients of the first step, before any update
    loss = loss_of(model, x, y)
    opt.zero_grad()
    dp.backward(loss)
    grads = [p.grad.numpy().copy() for p in model.parameters()]
    opt.step()
    # (2) a few whole steps through the path bench.py uses (exchange overlapped with backward, then / fused with Adam)
    losses = [float(loss.item())]
    for _ in range(STEPS - 1):
        loss = loss_of(model, x, y)
        opt.zero_grad()
        dp.backward_and_step(loss)
        losses.append(float(loss.item()))
    params = [p.numpy().copy() for p in model.parameters()]
    rt.synchronize()
    comm.barrier()
    result = None
    if rank == 0:
        ref_model, ref_opt = build()
        xf = CudaTensor.from_numpy(ids, requires_grad=False)
        yf = CudaTensor.from_numpy(labels, requires_grad=False)
        ref_losses, ref_grads = [], None
        for s in range(STEPS):
            ref_loss = loss_of(ref_model, xf, yf)
            ref_opt.zero_grad()
            ref_loss.backward()
            if s == 0:
                ref_grads = [p.grad.numpy().copy() for p in ref_model.parameters()]
            ref_opt.step()
            ref_losses.append(float(ref_loss.item()))
        ref_params = [p.numpy().copy() for p in ref_model.parameters()]
        gmax = max(float(np.abs(g).max()) for g in ref_grads)
        g_err = max(float(np.abs(g - r).max()) for g, r in zip(grads, ref_grads)) / gmax
        pmax = max(float(np.abs(p).max()) for p in ref_params)
        p_err = max(float(np.abs(p - r).max()) for p, r in zip(params, ref_params)) / pmax
        result = {'world': world, 'mode': mode, 'exchange': getattr(dp, 'exchange', 'nccl'),
                  'exchange_note': getattr(dp, 'exchange_note', ''), 'global_batch': BATCH,
                  'grad_rel_err': g_err, 'param_rel_err_after_%d_steps' % STEPS: p_err,
                  'local_losses': losses, 'global_losses': ref_losses}
    # every rank holds the same parameters after the exchange
    flat = np.concatenate([p.reshape(-1) for p in params]).astype(np.float64)
    checksum = float(np.abs(flat).sum())
    sums = [comm.max_float(checksum), -comm.max_float(-checksum)]
    if rank == 0:
        result['replica_checksum_spread'] = sums[0] - sums[1]
        with open(out_path, 'w') as f:
            json.dump(result, f)
    dp.close()
    comm.close()


if __name__ == '__main__':
    main()
