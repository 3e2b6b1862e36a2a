"""Data-parallel host logic on CPU: 2 ranks over gloo, batch sharded, gradients averaged.
Checks SURVEY.md 8(d) config 5's parity rule: N-rank averaged gradients == 1-rank gradients on the
concatenated batch (the loss is a mean over rows, local batches are equal)."""
import os
import subprocess
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, %(root)r)
import lightgrad_b200 as light
import lightgrad_b200.nn as nn
from lightgrad_b200 import parallel
from oracle import CpuTensor
from examples import mnist as mn

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
comm = parallel.GlooComm()
with nn.use_tensor(CpuTensor):
    np.random.seed(100 + rank)            # deliberately different: the wrapper must broadcast rank 0's
    model = mn.NN()
opt = light.optim.SGD(model.parameters(), lr=1e-3)
dp = parallel.DataParallel(model, opt, comm=comm)
x, y = mn.synthetic_batch(batch=16, seed=5)
lo, hi = dp.shard(16)
xl = CpuTensor.from_numpy(x[lo:hi], requires_grad=False)
yl = CpuTensor.from_numpy(y[lo:hi], requires_grad=False)
logits = model(xl)
loss = light.loss.cross_entropy(logits, yl)
opt.zero_grad()
loss.backward()
dp.sync_gradients()
np.savez(os.path.join(%(out)r, 'rank%%d.npz' %% rank), w1=model.l1.weight.numpy(), g1=model.l1.weight.grad.numpy(),
         g2=model.l2.weight.grad.numpy(), lo=lo, hi=hi)
'''


def test_two_rank_gradients_equal_single_rank(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT, out=str(tmp_path)))
    port = 29500 + (os.getpid() % 500)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                   LOCAL_RANK=str(r), OMP_NUM_THREADS='1')
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    r0, r1 = np.load(tmp_path / 'rank0.npz'), np.load(tmp_path / 'rank1.npz')
    assert (int(r0['lo']), int(r0['hi']), int(r1['lo']), int(r1['hi'])) == (0, 8, 8, 16)
    np.testing.assert_array_equal(r0['w1'], r1['w1'])           # broadcast made the replicas identical
    np.testing.assert_array_equal(r0['g1'], r1['g1'])           # all-reduce left identical gradients
    # single-process reference on the whole batch with rank 0's initial parameters
    sys.path.insert(0, ROOT)
    import lightgrad_b200 as light
    import lightgrad_b200.nn as nn
    from oracle import CpuTensor
    from examples import mnist as mn
    with nn.use_tensor(CpuTensor):
        np.random.seed(100)
        model = mn.NN()
    x, y = mn.synthetic_batch(batch=16, seed=5)
    loss = light.loss.cross_entropy(model(CpuTensor.from_numpy(x, requires_grad=False)),
                                    CpuTensor.from_numpy(y, requires_grad=False))
    for p in model.parameters():
        p.zero_grad()
    loss.backward()
    np.testing.assert_allclose(r0['g1'], model.l1.weight.grad.numpy(), rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(r0['g2'], model.l2.weight.grad.numpy(), rtol=1e-5, atol=1e-8)


def test_shard_rows():
    from lightgrad_b200.parallel import shard_rows
    assert [shard_rows(256, r, 8) for r in (0, 7)] == [(0, 32), (224, 256)]
    assert shard_rows(256, 1, 2) == (128, 256)


def test_pipelined_bucket_step_equals_plain_step(fake_device, monkeypatch):
    """DataParallel.backward_and_step (per-bucket all-reduce + Adam behind it) against loss.backward();
    optimizer.step() on the numpy stand-in of the device (its all-reduce is the identity): same parameters,
    Adam state and step counter after several steps, for bucket sizes from one parameter to everything."""
    import lightgrad_b200 as light
    import lightgrad_b200.nn as nn
    from lightgrad_b200 import CudaTensor, parallel
    from examples import bert

    monkeypatch.setenv('LG_DP_PIPELINED_STEP', '1')
    cfg = dict(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=2, vocab_size=50,
               max_position_embeddings=16, type_vocab_size=2)
    ids, labels = bert.synthetic_batch(2, 8, cfg['vocab_size'])

    def build():
        with nn.use_tensor(CudaTensor):
            np.random.seed(3)
            model = bert.BertForMaskedLM(**cfg)
        return model, light.optim.Adam(model.parameters(), lr=1e-2)

    def run(bucket_bytes):
        model, opt = build()
        dp = None
        if bucket_bytes is not None:
            dp = parallel.DataParallel(model, opt, comm=parallel.LocalComm())
            dp.world, dp._nccl = 2, True             # take the multi-GPU code path; the fake all-reduce is the identity
        x = CudaTensor.from_numpy(ids, requires_grad=False)
        y = CudaTensor.from_numpy(labels, requires_grad=False)
        for _ in range(3):
            loss = light.loss.cross_entropy(model(x).reshape(-1, cfg['vocab_size']), y)
            opt.zero_grad()
            if dp is None:
                loss.backward()
                opt.step()
            else:
                dp.backward_and_step(loss, bucket_bytes=bucket_bytes)
        return [p.numpy() for p in model.parameters()], opt.t, loss.item()

    want, t_want, loss_want = run(None)
    for bucket_bytes in (1, 4096, 1 << 30):
        got, t_got, loss_got = run(bucket_bytes)
        assert t_got == t_want
        assert loss_got == loss_want
        for a, b in zip(got, want):
            np.testing.assert_array_equal(a, b)
