"""Whole-step CUDA graph capture: replaying a captured training step must train exactly like eager steps."""
import numpy as np
import pytest
import lightgrad_b200 as light
import lightgrad_b200.nn as nn
from lightgrad_b200 import CudaTensor


def _make(seed=0):
    from examples import bert
    cfg = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, vocab_size=300,
               max_position_embeddings=32, type_vocab_size=2)
    np.random.seed(seed)
    model = bert.BertForMaskedLM(**cfg)
    model.load_parameters({n: (p.numpy() * 8).astype(np.float32) if 'LayerNorm' not in n and not n.endswith('predictions.bias')
                           else p.numpy() for n, p in model.named_parameters()})
    opt = light.optim.Adam(model.parameters(), lr=1e-3)
    ids, labels = bert.synthetic_batch(4, 16, cfg['vocab_size'])
    return bert, model, opt, CudaTensor.from_numpy(ids, requires_grad=False), CudaTensor.from_numpy(labels, requires_grad=False)


@pytest.mark.gpu
def test_replayed_steps_match_eager_steps(cuda):
    from lightgrad_b200.autograd.cuda.graph import StepGraph
    bert, m1, o1, ids, labels = _make()
    eager_losses = [bert.train_step(m1, o1, ids, labels).item() for _ in range(5)]
    bert, m2, o2, ids2, labels2 = _make()
    sg = StepGraph(lambda: bert.train_step(m2, o2, ids2, labels2), warmup=2)   # steps 1, 2 eager; step 3 = first launch
    losses = [float(sg.outputs.item())]
    for _ in range(2):
        losses.append(float(sg.replay().item()))
    np.testing.assert_allclose(losses, eager_losses[2:], rtol=2e-5)
    for (n, p), (_, q) in zip(m1.named_parameters(), m2.named_parameters()):
        # atomics (split-K reduce-add, scatter-add, LN partials) reorder sums; Adam turns that into <= a few 1e-6
        np.testing.assert_allclose(p.numpy(), q.numpy(), rtol=1e-4, atol=5e-5, err_msg=n)
    assert sg.n_kernels > 50 and sg.n_nodes >= sg.n_kernels


@pytest.mark.gpu
def test_host_round_trips_are_rejected_during_capture(cuda):
    from lightgrad_b200.autograd.cuda.graph import StepGraph
    x = CudaTensor.from_numpy(np.ones((4, 4), dtype=np.float32))
    with pytest.raises(RuntimeError):
        StepGraph(lambda: (x * 2).sum().item() if cuda.api.raw and _capturing(cuda) else (x * 2).sum(), warmup=0)
    # the library is usable again afterwards
    assert float((x * 2).sum().item()) == 32.0


def _capturing(cuda):
    return True
