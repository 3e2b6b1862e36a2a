"""Where does the bf16 mode spend its error budget?  One full-size BERT-base step (batch B, seq 128) on the CPU oracle
and on the device with different matmul classes in bf16 (the rest in tf32): worst per-tensor gradient error relative
to that tensor's largest gradient, and the tensors that carry it.

    python benchmarks/bf16_error_study.py [--batch 2] [--configs FXW,FXWA,FW,XW,W,F,X,]
"""
import argparse
import json
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_full_size import _step_grads, _worst_rel_err            # noqa: E402
from lightgrad_b200 import CudaTensor                                   # noqa: E402
from lightgrad_b200.autograd.cuda import ops                            # noqa: E402
from examples import bert                                               # noqa: E402
from oracle import CpuTensor                                            # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=2)
    ap.add_argument('--configs', default='FXW,FXWA,FW,XW,W,F,X,')
    a = ap.parse_args()
    cfg = dict(bert.BERT_BASE)
    l_ref, g_ref = _step_grads(CpuTensor, cfg, a.batch, 128)
    gmax = max(float(np.abs(g).max()) for g in g_ref.values())
    for mode, classes in [('tf32', None)] + [('bf16', c) for c in a.configs.split(',')]:
        ops.set_matmul_mode(mode)
        if classes is not None:
            ops.set_bf16_classes(classes)
        l, g = _step_grads(CudaTensor, cfg, a.batch, 128)
        worst, where = _worst_rel_err(g_ref, g)
        errs = sorted(((float(np.abs(g_ref[n] - g[n]).max() / max(float(np.abs(g_ref[n]).max()), 1e-6 * gmax)), n)
                       for n in g_ref), reverse=True)
        fro = float(np.sqrt(sum(((g_ref[n] - g[n]) ** 2).sum() for n in g_ref)) /
                    np.sqrt(sum((g_ref[n] ** 2).sum() for n in g_ref)))
        print(json.dumps({'mode': mode, 'bf16_classes': classes, 'batch': a.batch, 'loss_rel_err': abs(l - l_ref) / abs(l_ref),
                          'worst_rel_err': worst, 'where': where, 'global_frobenius_rel_err': fro,
                          'median_tensor_err': errs[len(errs) // 2][0], 'top5': errs[:5]}), flush=True)


if __name__ == '__main__':
    main()
