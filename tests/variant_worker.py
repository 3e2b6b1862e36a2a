"""One tf32 training step of a small BERT on the GPU; prints the loss and a checksum of every parameter gradient as
JSON.  Run by tests/test_variants.py in a subprocess with one of the opt-in kernel variants switched on through its
environment variable (the variables are read once per process)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                                                      # noqa: E402
import lightgrad_b200 as light                                          # noqa: E402
import lightgrad_b200.nn                                                # noqa: E402,F401
from lightgrad_b200 import CudaTensor                                   # noqa: E402
from lightgrad_b200.autograd.cuda import ops                            # noqa: E402
from examples import bert                                               # noqa: E402

ops.set_matmul_mode('tf32')
np.random.seed(7)
cfg = dict(bert.BERT_BASE, hidden_size=768, intermediate_size=3072, num_hidden_layers=2, num_attention_heads=12,
           vocab_size=4096, max_position_embeddings=128)
with light.nn.use_tensor(CudaTensor):
    model = bert.BertForMaskedLM(**cfg)
ids, labels = bert.synthetic_batch(4, 128, cfg['vocab_size'])
x, y = CudaTensor.from_numpy(ids, requires_grad=False), CudaTensor.from_numpy(labels, requires_grad=False)
opt = light.optim.Adam(model.parameters(), lr=1e-3)
loss = light.loss.cross_entropy(model(x).reshape(-1, cfg['vocab_size']), y)
opt.zero_grad()
loss.backward()
grads = [p.grad.numpy().astype(np.float64) for p in model.parameters()]
print(json.dumps({'loss': float(loss.item()), 'norms': [float(np.sqrt((g * g).sum())) for g in grads],
                  'sums': [float(g.sum()) for g in grads]}))
