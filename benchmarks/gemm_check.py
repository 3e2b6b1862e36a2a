"""Correctness sweep of lg_gemm in a tensor-core mode against float64 numpy (diagnostic)."""
import os, sys, itertools
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lightgrad_b200 as light
from lightgrad_b200 import CudaTensor as T
from lightgrad_b200.autograd.cuda import ops, runtime as rt

mode = sys.argv[1] if len(sys.argv) > 1 else 'tf32'
ops.set_matmul_mode(mode)
rs = np.random.RandomState(0)
shapes = [(128, 256, 32), (128, 256, 64), (128, 64, 256), (256, 512, 256), (384, 192, 96), (4096, 768, 768), (300, 520, 136),
          (768, 768, 4096), (4096, 3072, 768), (1024, 30522 // 8 * 8 + 4, 256), (130, 70, 40), (2048, 2048, 2048)]
bad = 0
with light.no_grad():
    for (M, N, K) in shapes:
        for ta, tb in itertools.product((False, True), repeat=2):
            a = rs.uniform(-1, 1, (K, M) if ta else (M, K)).astype(np.float32)
            b = rs.uniform(-1, 1, (N, K) if tb else (K, N)).astype(np.float32)
            A, B = T.from_numpy(a), T.from_numpy(b)
            Av = A.transpose(1, 0) if ta else A
            Bv = B.transpose(1, 0) if tb else B
            for with_bias in (False, True):
                bias = rs.uniform(-1, 1, N).astype(np.float32) if with_bias else None
                out = ops._gemm(Av, Bv, bias=T.from_numpy(bias) if with_bias else None)
                got = out.numpy()
                want = (a.T if ta else a).astype(np.float64) @ (b.T if tb else b).astype(np.float64)
                if with_bias:
                    want = want + bias
                err = np.abs(got - want).max() / (np.abs(want).max() + 1e-30)
                ok = err < 5e-3 and np.isfinite(got).all()
                bad += (not ok)
                print("%s M=%5d N=%5d K=%5d ta=%d tb=%d bias=%d  rel_err=%.2e  shape=%s" %
                      ("ok " if ok else "BAD", M, N, K, ta, tb, with_bias, err, got.shape), flush=True)
# batched problems with strided (head-split) views, as BertSelfAttention issues them
with light.no_grad():
    b, h, sq, d = 4, 12, 128, 64
    q = rs.uniform(-1, 1, (b, sq, h * d)).astype(np.float32)
    k = rs.uniform(-1, 1, (b, sq, h * d)).astype(np.float32)
    v = rs.uniform(-1, 1, (b, sq, h * d)).astype(np.float32)
    Q = T.from_numpy(q).reshape(b, sq, h, d).transpose(0, 2, 1, 3)
    Kt = T.from_numpy(k).reshape(b, sq, h, d).transpose(0, 2, 3, 1)
    V = T.from_numpy(v).reshape(b, sq, h, d).transpose(0, 2, 1, 3)
    qn = q.reshape(b, sq, h, d).transpose(0, 2, 1, 3).astype(np.float64)
    kn = k.reshape(b, sq, h, d).transpose(0, 2, 3, 1).astype(np.float64)
    vn = v.reshape(b, sq, h, d).transpose(0, 2, 1, 3).astype(np.float64)
    p_ = rs.uniform(0, 1, (b, h, sq, sq)).astype(np.float32)
    P = T.from_numpy(p_)
    cases = [("QK^T", Q, Kt, qn @ kn), ("PV", P, V, p_.astype(np.float64) @ vn),
             ("dP=dC V^T", T.from_numpy(q).reshape(b, sq, h, d).transpose(0, 2, 1, 3), ops._swap_last(V), qn @ vn.transpose(0, 1, 3, 2)),
             ("dV=P^T dC", ops._swap_last(P), Q, p_.astype(np.float64).transpose(0, 1, 3, 2) @ qn),
             ("dK^T=Q^T dS", ops._swap_last(Q), P, qn.transpose(0, 1, 3, 2) @ p_.astype(np.float64)),
             ("3d batch", T.from_numpy(q), T.from_numpy(k).transpose(0, 2, 1), q.astype(np.float64) @ k.astype(np.float64).transpose(0, 2, 1))]
    for name, A, B, want in cases:
        got = ops._gemm(A, B).numpy()
        err = np.abs(got - want).max() / (np.abs(want).max() + 1e-30)
        ok = err < 5e-3 and np.isfinite(got).all()
        bad += (not ok)
        print("%s batched %-12s rel_err=%.2e shape=%s" % ("ok " if ok else "BAD", name, err, got.shape), flush=True)
print("bad:", bad)
