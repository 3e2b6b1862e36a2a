"""One BERT-base training step inside a cudaProfilerStart/Stop range, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --csv ... python benchmarks/profile_step.py
(only the kernels of the bracketed step are instrumented, so the capture is quick)."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lightgrad_b200 as light
from lightgrad_b200 import CudaTensor
from lightgrad_b200.autograd.cuda import runtime as rt, ops
from examples import bert
import bench

ap = argparse.ArgumentParser()
ap.add_argument('--mode', default='tf32')
ap.add_argument('--batch', type=int, default=32)
ap.add_argument('--warmup', type=int, default=2)
ap.add_argument('--steps', type=int, default=1)
ap.add_argument('--layers', type=int, default=None, help="encoder layers (default: BERT-base's 12); 1 gives one launch of "
                "each kernel at the step's shapes for an `ncu --set full` capture")
args = ap.parse_args()
ops.set_matmul_mode(args.mode)
cfg = dict(bert.BERT_BASE)
if args.layers:
    cfg['num_hidden_layers'] = args.layers
model = bench.build_bert(CudaTensor, cfg)
opt = light.optim.Adam(model.parameters(), lr=1e-4)
step = bench.make_step(model, opt, None, light)
light.Gradients.retain_intermediate = False
ids, labels = bert.synthetic_batch(args.batch, 128, bert.BERT_BASE['vocab_size'])
ids_d, lab_d = CudaTensor.from_numpy(ids, requires_grad=False), CudaTensor.from_numpy(labels, requires_grad=False)
for _ in range(args.warmup):
    step(ids_d, lab_d)
rt.synchronize()
n0 = rt.launch_count()
rt.api.profiler_range(1)
for _ in range(args.steps):
    loss = step(ids_d, lab_d)
rt.api.profiler_range(0)
print("launches in range:", rt.launch_count() - n0, "loss", loss.item())
