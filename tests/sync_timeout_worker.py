"""Worker of tests/test_parallel_gpu.py::test_stuck_peer_fails_fast: two ranks set up the NCCL communicator; rank 0 then
issues an all-reduce that rank 1 never joins (rank 1 just sleeps and leaves).  With LG_SYNC_TIMEOUT_S set, rank 0's
synchronisation must come back with an error that names the cause instead of waiting for ever, after which the process
is still usable (the communicator was aborted, its kernel has left the GPU)."""
import json
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lightgrad_b200 import CudaTensor, parallel                        # noqa: E402
from lightgrad_b200.autograd.cuda import runtime as rt                 # noqa: E402


def main():
    out_path = sys.argv[1]
    rt.ensure_device(int(os.environ.get('LOCAL_RANK', '0')))
    comm = parallel.default_comm()
    rank = comm.rank
    ident = None
    if rank == 0:
        buf = np.zeros(128, dtype=np.uint8)
        rt.api.nccl_unique_id(buf.ctypes.data)
        ident = buf.tobytes()
    ident = comm.broadcast_bytes(ident, root=0)
    idbuf = np.frombuffer(ident, dtype=np.uint8).copy()
    rt.api.nccl_init(idbuf.ctypes.data, comm.world, rank)
    t = CudaTensor.from_numpy(np.ones(1 << 20, dtype=np.float32), requires_grad=False)
    rt.api.nccl_allreduce_f32(t.ptr, 1 << 20, 0, 0)         # both ranks: the communicator works
    rt.synchronize()
    comm.barrier()
    if rank == 1:
        time.sleep(12)                                       # never joins the second collective
        comm.barrier()
        return
    rt.api.nccl_allreduce_f32(t.ptr, 1 << 20, 0, 0)
    t0 = time.time()
    try:
        rt.synchronize()
        outcome = {'raised': False}
    except RuntimeError as exc:
        outcome = {'raised': True, 'message': str(exc), 'seconds': time.time() - t0}
    # the process is still usable: a plain kernel and a read-back
    u = CudaTensor.from_numpy(np.arange(8, dtype=np.float32), requires_grad=False)
    outcome['after'] = float((u + u).numpy().sum())
    json.dump(outcome, open(out_path, 'w'))
    comm.barrier()


if __name__ == '__main__':
    main()
