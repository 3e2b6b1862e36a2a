"""Host-side cost of one BERT-base step: enqueue time vs device time, and a cProfile of the python dispatch."""
import cProfile, pstats, io, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lightgrad_b200 as light
from lightgrad_b200 import CudaTensor
from lightgrad_b200.autograd.cuda import runtime as rt, ops
from examples import bert
import bench

ops.set_matmul_mode('tf32')
model = bench.build_bert(CudaTensor, bert.BERT_BASE)
opt = light.optim.Adam(model.parameters(), lr=1e-4)
step = bench.make_step(model, opt, None, light)
light.Gradients.retain_intermediate = False
ids, labels = bert.synthetic_batch(32, 128, bert.BERT_BASE['vocab_size'])
ids_d, lab_d = CudaTensor.from_numpy(ids, requires_grad=False), CudaTensor.from_numpy(labels, requires_grad=False)
for _ in range(3):
    step(ids_d, lab_d)
rt.synchronize()
K = 10
t0 = time.perf_counter()
for _ in range(K):
    step(ids_d, lab_d)
t1 = time.perf_counter()
rt.synchronize()
t2 = time.perf_counter()
print("enqueue %.2f ms/step, total %.2f ms/step" % ((t1 - t0) / K * 1e3, (t2 - t0) / K * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step(ids_d, lab_d)
pr.disable()
rt.synchronize()
out = io.StringIO()
pstats.Stats(pr, stream=out).sort_stats('tottime').print_stats(35)
print(out.getvalue()[:6000])
