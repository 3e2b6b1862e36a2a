"""Prints every golden case where CudaTensor deviates (diagnostic; the pass/fail gate is tests/)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightgrad_b200 import CudaTensor
from tests import replay

def err(got, want):
    d = np.abs(got.astype(np.float64) - want.astype(np.float64))
    return float(d.max()) if d.size else 0.0, float(d.max() / (np.abs(want).max() + 1e-30)) if d.size else 0.0

bad = 0
it = replay.replay_ops(CudaTensor)
while True:
    try:
        case, field, got, want = next(it)
    except StopIteration:
        break
    except Exception as e:
        print("EXC in replay_ops:", repr(e)); break
    if got.shape != want.shape:
        print("SHAPE", case, field, got.shape, want.shape); bad += 1; continue
    a, r = err(got, want)
    if r > 5e-6:
        print("BAD %-28s %-10s abs %.3e rel %.3e" % (case, field, a, r)); bad += 1
print("ops: %d bad" % bad)
try:
    for name, got, want in replay.replay_bert_tiny(CudaTensor):
        a, r = err(got, want)
        print("%s %-60s abs %.3e rel %.3e" % ("BAD" if r > 1e-4 else "ok ", name, a, r))
except Exception as e:
    import traceback; traceback.print_exc()
