"""Data parallelism on real GPUs (SURVEY.md 8(d) config 5 parity rule): two ranks, one process per GPU, NCCL over
NVLink.  The all-reduced gradients of a sharded batch equal the single-process gradients of the whole batch, the
replicas hold identical parameters after several steps, and every rank leaves through a normal interpreter exit.
Skipped on a box with fewer than two GPUs (the CPU suite covers the host logic over gloo, tests/test_parallel.py)."""
import json
import os
import subprocess
import sys
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import ctypes
    from lightgrad_b200.autograd.cuda import runtime as rt
    n = ctypes.c_int(0)
    rt.load().raw.lg_device_count(ctypes.byref(n))
    return n.value


def _run_two_ranks(tmp_path, mode, env_extra=None):
    out = tmp_path / ('dp_%s.json' % mode)
    port = 29600 + (os.getpid() % 300)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE='2', LOCAL_RANK=str(r), MASTER_ADDR='127.0.0.1',
                   MASTER_PORT=str(port))
        env.update(env_extra or {})
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, 'tests', 'dp_worker.py'), str(out), mode],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    logs = []
    for p in procs:
        try:
            logs.append(p.communicate(timeout=600)[0].decode())
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    return json.load(open(out))


@pytest.mark.parametrize('exchange', ['nccl', 'nvls'])
@pytest.mark.parametrize('mode,tol', [('fp32', 1e-5), ('tf32', 5e-3)])
def test_two_gpu_gradients_equal_single_gpu(tmp_path, mode, tol, exchange):
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs")
    res = _run_two_ranks(tmp_path, mode, {'LG_DP_EXCHANGE': exchange})
    assert res['world'] == 2
    if res['exchange'] != exchange:
        pytest.skip("asked for the %s exchange, ran %s: %s" % (exchange, res['exchange'], res.get('exchange_note')))
    assert res['grad_rel_err'] <= tol, res
    # lr = 1e-3 Adam steps move every element by ~lr per step whatever the gradient's size, so parameters are
    # compared loosely: a sign flip of a ~0 gradient element is 2 lr
    assert res['param_rel_err_after_3_steps'] <= 0.05, res
    assert res['replica_checksum_spread'] == 0.0, res
    assert abs(res['local_losses'][0] - res['global_losses'][0]) < 0.05


def test_stuck_peer_fails_fast(tmp_path):
    """A collective whose peer never arrives must not hang the rank: the watched synchronisation
    aborts the communicator after LG_SYNC_TIMEOUT_S and raises; the process stays usable."""
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs")
    out = tmp_path / 'stuck.json'
    port = 29300 + (os.getpid() % 300)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE='2', LOCAL_RANK=str(r), MASTER_ADDR='127.0.0.1',
                   MASTER_PORT=str(port), LG_SYNC_TIMEOUT_S='4')
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, 'tests', 'sync_timeout_worker.py'), str(out)],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    logs = []
    for p in procs:
        try:
            logs.append(p.communicate(timeout=120)[0].decode())
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    res = json.load(open(out))
    assert res['raised'], res
    assert 'peer rank' in res['message'] and 'LG_SYNC_TIMEOUT_S' in res['message'], res
    assert 3.5 <= res['seconds'] <= 10.0, res
    assert res['after'] == 56.0
