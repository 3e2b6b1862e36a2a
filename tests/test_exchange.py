"""Gradient exchange fused with the optimizer over NVLink multicast (csrc/lg_mc.cu, DataParallel exchange='nvls').

A team of ONE device exercises the whole path -- multicast object, region layout, arena migration, per-bucket
reduce-scatter / update / all-gather kernel with its cross-GPU flags, step counter -- and must leave exactly the
parameters that ``loss.backward(); optimizer.step()`` leaves (the reduce-scatter of one device is the identity).
Runs on the numpy test double in the CPU suite (host logic) and on the real device with `-m gpu`; the two-GPU
version is tests/test_parallel_gpu.py."""
import numpy as np
import pytest
import lightgrad_b200 as light
import lightgrad_b200.nn as nn
from lightgrad_b200 import CudaTensor, parallel
from examples import bert

CFG = dict(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=2, vocab_size=50,
           max_position_embeddings=16, type_vocab_size=2)


def _build(make_opt):
    with nn.use_tensor(CudaTensor):
        np.random.seed(3)
        model = bert.BertForMaskedLM(**CFG)
    return model, make_opt(model.parameters())


def _run(make_opt, bucket_bytes, steps=3, exchange='nvls'):
    ids, labels = bert.synthetic_batch(2, 8, CFG['vocab_size'])
    model, opt = _build(make_opt)
    dp = None
    if bucket_bytes is not None:
        dp = parallel.DataParallel(model, opt, comm=parallel.LocalComm(), exchange=exchange)
        if dp.exchange != exchange:
            pytest.skip("NVLink multicast objects are not available on this device: " + dp.exchange_note)
    x = CudaTensor.from_numpy(ids, requires_grad=False)
    y = CudaTensor.from_numpy(labels, requires_grad=False)
    for _ in range(steps):
        loss = light.loss.cross_entropy(model(x).reshape(-1, CFG['vocab_size']), y)
        opt.zero_grad()
        if dp is None:
            loss.backward()
            opt.step()
        else:
            dp.backward_and_step(loss, bucket_bytes=bucket_bytes)
    out = [p.numpy() for p in model.parameters()], getattr(opt, 't', None), loss.item()
    if dp is not None:
        dp.close()
    return out


OPTIMIZERS = {
    'adam': lambda ps: light.optim.Adam(ps, lr=1e-2),
    'adabelief': lambda ps: light.optim.AdaBelief(ps, lr=1e-2),
    'sgd': lambda ps: light.optim.SGD(ps, lr=1e-2),
    'sgd_momentum': lambda ps: light.optim.SGD(ps, lr=1e-2, momentum=0.9),
}


def _check(name, exchange='nvls', exact=True):
    """``exact``: bit-identical parameters (the numpy test double is deterministic).  On the device the gradients
    themselves are not bit-reproducible from run to run (atomic scatter-adds of the embedding gradient, split-K partial
    sums), so two runs of the PLAIN step differ in the last bits too: the comparison there is to 1e-4 / 1e-6."""
    want, t_want, loss_want = _run(OPTIMIZERS[name], None)
    for bucket_bytes in (1, 4096, 1 << 30):
        got, t_got, loss_got = _run(OPTIMIZERS[name], bucket_bytes, exchange=exchange)
        assert t_got == t_want
        if exact:
            assert loss_got == loss_want
        else:
            assert abs(loss_got - loss_want) <= 1e-5 * abs(loss_want)
        for a, b in zip(got, want):
            if exact:
                np.testing.assert_array_equal(a, b)
            else:
                np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize('name', sorted(OPTIMIZERS))
def test_team_of_one_exchange_equals_plain_step_host_logic(fake_device, name):
    _check(name)


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(OPTIMIZERS))
def test_team_of_one_exchange_equals_plain_step_on_device(cuda, name):
    _check(name, exact=False)


# one GPU: the optimizer as per-bucket small-footprint kernels on the collective stream, beside backward (exchange 'local')
@pytest.mark.parametrize('name', sorted(OPTIMIZERS))
def test_overlapped_optimizer_equals_plain_step_host_logic(fake_device, name):
    _check(name, exchange='local')


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(OPTIMIZERS))
def test_overlapped_optimizer_equals_plain_step_on_device(cuda, name):
    _check(name, exchange='local', exact=False)


@pytest.mark.gpu
def test_overlapped_optimizer_inside_a_captured_graph(cuda):
    from lightgrad_b200.autograd.cuda.graph import StepGraph
    ids, labels = bert.synthetic_batch(2, 8, CFG['vocab_size'])
    x = CudaTensor.from_numpy(ids, requires_grad=False)
    y = CudaTensor.from_numpy(labels, requires_grad=False)

    def run(graph):
        model, opt = _build(OPTIMIZERS['adam'])
        dp = parallel.DataParallel(model, opt, comm=parallel.LocalComm(), exchange='local')
        assert dp.exchange == 'local'

        def step():
            loss = light.loss.cross_entropy(model(x).reshape(-1, CFG['vocab_size']), y)
            opt.zero_grad()
            dp.backward_and_step(loss, bucket_bytes=4096)
            return loss
        if graph:
            sg = StepGraph(step, warmup=1)          # one eager step (shape-dependent constants), then capture = step 2
            for _ in range(3):
                sg.replay()
            sg.destroy()
        else:
            for _ in range(5):
                step()
        return [p.numpy() for p in model.parameters()]
    # five Adam steps: for an element whose gradient is near zero, m / sqrt(v) turns last-bit differences of the gradient
    # (atomic scatter-adds, split-K) into a visible fraction of lr per step; measured 1.4e-6 on one element in 1024
    for a, b in zip(run(True), run(False)):
        np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
def test_exchange_step_inside_a_captured_graph(cuda):
    """The exchange kernels keep their cross-GPU epochs on the device, so a captured step replays correctly."""
    from lightgrad_b200.autograd.cuda.graph import StepGraph
    ids, labels = bert.synthetic_batch(2, 8, CFG['vocab_size'])
    x = CudaTensor.from_numpy(ids, requires_grad=False)
    y = CudaTensor.from_numpy(labels, requires_grad=False)

    def run(graph):
        model, opt = _build(OPTIMIZERS['adam'])
        dp = parallel.DataParallel(model, opt, comm=parallel.LocalComm(), exchange='nvls')
        if dp.exchange != 'nvls':
            pytest.skip("NVLink multicast objects are not available on this device")

        def step():
            loss = light.loss.cross_entropy(model(x).reshape(-1, CFG['vocab_size']), y)
            opt.zero_grad()
            dp.backward_and_step(loss, bucket_bytes=4096)
            return loss
        if graph:
            sg = StepGraph(step, warmup=1)          # one eager step, then the capture runs the step once
            for _ in range(3):
                sg.replay()
            sg.destroy()
        else:
            for _ in range(5):
                step()
        out = [p.numpy() for p in model.parameters()]
        dp.close()
        return out
    # five Adam steps: for an element whose gradient is near zero, m / sqrt(v) turns last-bit differences of the gradient
    # (atomic scatter-adds, split-K) into a visible fraction of lr per step; measured 1.4e-6 on one element in 1024
    for a, b in zip(run(True), run(False)):
        np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-5)
