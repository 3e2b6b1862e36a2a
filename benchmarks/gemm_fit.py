"""Device time of lg_gemm (tf32) vs K at fixed M, N, measured by replaying a CUDA graph of 40 back-to-back
calls (no host dispatch in the timed region): separates fixed per-launch cost from per-k-block cost."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lightgrad_b200 as light
from lightgrad_b200 import CudaTensor as T
from lightgrad_b200.autograd.cuda import ops, runtime as rt
from lightgrad_b200.autograd.cuda.graph import StepGraph

ops.set_matmul_mode('tf32')
rs = np.random.RandomState(0)
REP = 40
SHAPES = [tuple(int(v) for v in a.split('x')) for a in sys.argv[1:]] or \
    [(M, N, K) for (M, N) in ((4096, 768), (4096, 3072), (768, 768), (4096, 2304)) for K in (32, 128, 256, 768, 1536, 3072, 4096)]
for (M, N, K) in SHAPES:
    if True:
        a = T.from_numpy(rs.uniform(-1, 1, (M, K)).astype(np.float32))
        w = T.from_numpy(rs.uniform(-1, 1, (N, K)).astype(np.float32))
        outs = []

        def body():
            with light.no_grad():
                for _ in range(REP):
                    o = ops._gemm(a, ops._swap_last(w))
            return o
        sg = StepGraph(body, warmup=1)
        for _ in range(3):
            sg.replay()
        rt.synchronize()
        e0 = rt.Event().record()
        for _ in range(5):
            sg.replay()
        e1 = rt.Event().record()
        e1.synchronize()
        us = e0.elapsed_ms(e1) * 1e3 / (5 * REP)
        print(json.dumps(dict(M=M, N=N, K=K, us=round(us, 2), tflops=round(2.0 * M * N * K / us / 1e6, 1))), flush=True)
        del sg, a, w
