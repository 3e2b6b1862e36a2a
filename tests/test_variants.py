"""The opt-in kernel variants compute the same step as the defaults: a 2-layer BERT at the
benchmark's layer shapes (d 768, 12 heads, seq 128, so the tcgen05 GEMMs, the fused attention and the vectorised
LayerNorm all run), loss and every parameter gradient's norm compared with the default build of the same process
image.  Each variant needs its own process: the switches are read once."""
import json
import os
import subprocess
import sys
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'variant_worker.py')], env=env,
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    return json.loads(out.stdout.decode().strip().splitlines()[-1])


@pytest.fixture(scope='module')
def default_step(cuda):
    return _run({})


@pytest.mark.parametrize('env', [{'LG_GEMM_DYNAMIC': '1'}, {'LG_GEMM_PAIR': '0'}, {'LG_LN_ATOMIC': '1'},
                                 {'LG_NO_FUSED_ATTENTION': '1'}, {'LG_NO_PDL_SMALL': '1', 'LG_GEMM_NO_PDL': '1'},
                                 {'LG_PREFER_SHARED': '1', 'LG_GEMM_CARVEOUT': '1'}],
                         ids=lambda e: '+'.join(sorted(e)))
def test_variant_matches_default(cuda, default_step, env):
    got = _run(env)
    # tf32 products in a different summation order (tile shapes, split-K, atomics): 1e-3 of each gradient's norm
    assert abs(got['loss'] - default_step['loss']) <= 1e-4 * abs(default_step['loss'])
    for a, b in zip(got['norms'], default_step['norms']):
        assert abs(a - b) <= 2e-3 * max(b, 1e-6), (a, b)
