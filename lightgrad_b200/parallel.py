"""Data-parallel training: one process per GPU, batch sharded across ranks, gradients averaged.

The reference has no multi-device path at all (operands of different tensor classes are rejected,
lightgrad/autograd/func.py:18-20).  Only the batch axis of examples/mnist.py and
examples/bert.py shards naturally, so the exchange step is exactly one collective per iteration: an
averaging all-reduce of the parameter gradients, between ``loss.backward()`` and ``optimizer.step()``.

  * cuda backend: the optimizer keeps all gradients in one flat fp32 arena (optim._Arena).  Default exchange
    ("nvls"): both arenas move into an NVLink multicast region and every bucket of gradients is reduce-scattered
    through the switch, updated by the optimizer on the owning rank's 1/N share and all-gathered as new parameters
    by ONE kernel that runs next to the GEMMs of backward (C-ABI lg_mc_*, csrc/lg_mc.cu).  Fallback / explicit
    ``LG_DP_EXCHANGE=nccl``: in-place ncclAllReduce(avg) per bucket (C-ABI lg_nccl_*), then the optimizer.
  * any other backend (the CPU oracle in the gloo tests): gradients are averaged tensor by tensor
    through ``comm.allreduce_avg_numpy``.

Bootstrap (exchanging the 128-byte NCCL id, barriers, max-over-ranks of timings) is control-plane
only and uses ``torch.distributed`` with the gloo backend when it is available, as launched by
``python -m torch.distributed.run`` (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT from the env).
"""
import os
import numpy as np


class LocalComm(object):
    """World of one: every collective is the identity."""
    rank, world = 0, 1

    def barrier(self):
        pass

    def allreduce_avg_numpy(self, a):
        return a

    def max_float(self, x):
        return float(x)

    def broadcast_bytes(self, b, root=0):
        return b

    def close(self):
        pass


class GlooComm(object):
    """Control-plane communicator over torch.distributed (gloo).  Host memory only."""

    def __init__(self, rank=None, world=None, init_method=None):
        import torch.distributed as dist
        self.dist = dist
        if not dist.is_initialized():
            kw = {}
            if rank is not None:
                kw = dict(rank=rank, world_size=world)
            if init_method is not None:
                kw['init_method'] = init_method
            dist.init_process_group(backend='gloo', **kw)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def barrier(self):
        self.dist.barrier()

    def allreduce_avg_numpy(self, a):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a).copy())
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return (t.numpy() / self.world).astype(a.dtype)

    def max_float(self, x):
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def broadcast_bytes(self, b, root=0):
        obj = [b if self.rank == root else None]
        self.dist.broadcast_object_list(obj, src=root)
        return obj[0]

    def close(self):
        """Leave the process group together (the last collective of the job)."""
        if self.dist.is_initialized():
            try:
                self.dist.barrier()
            finally:
                self.dist.destroy_process_group()


def default_comm():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world <= 1:
        return LocalComm()
    return GlooComm()


def _pass_fd(comm, fd):
    """Hand rank 0's file descriptor to every other rank of this node (SCM_RIGHTS over a unix socket).  Returns the
    local descriptor number on every rank."""
    import socket
    if comm.world == 1:
        return fd
    if comm.rank == 0:
        path = '/tmp/lg_mc_%d_%s.sock' % (os.getpid(), os.environ.get('MASTER_PORT', '0'))
        try:
            os.unlink(path)
        except OSError:
            pass
        srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        srv.bind(path)
        srv.listen(comm.world)
        comm.broadcast_bytes(path.encode(), root=0)
        try:
            srv.settimeout(120)
            for _ in range(comm.world - 1):
                conn, _addr = srv.accept()
                socket.send_fds(conn, [b'fd'], [fd])
                conn.recv(1)                       # the peer has the descriptor
                conn.close()
        finally:
            srv.close()
            os.unlink(path)
        return fd
    path = comm.broadcast_bytes(None, root=0).decode()
    c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    c.settimeout(120)
    c.connect(path)
    _msg, fds, _flags, _addr = socket.recv_fds(c, 16, 1)
    c.send(b'k')
    c.close()
    return fds[0]


def shard_rows(n_rows, rank, world):
    """Contiguous, equal slice [lo, hi) of the batch for ``rank``."""
    assert n_rows % world == 0, "global batch %d is not divisible by %d ranks" % (n_rows, world)
    per = n_rows // world
    return rank * per, (rank + 1) * per


class DataParallel(object):
    """Keeps replicas in step: identical initial parameters, averaged gradients.

        dp = DataParallel(model, optimizer)          # broadcasts rank 0's parameters
        x_local = x_global[slice(*dp.shard(len(x_global)))]
        loss = step_fn(...); loss.backward()
        dp.sync_gradients()                          # one averaging all-reduce
        optimizer.step()
    """

    def __init__(self, model, optimizer, comm=None, broadcast=True, exchange=None):
        """``exchange``: 'nvls' | 'nccl' | 'auto' (default: $LG_DP_EXCHANGE or 'auto' = nvls when the GPUs offer NVLink
        multicast and the optimizer has a fused exchange kernel, else nccl)."""
        self.model, self.optimizer = model, optimizer
        self.comm = comm if comm is not None else default_comm()
        self.rank, self.world = self.comm.rank, self.comm.world
        self.arena = getattr(optimizer, 'arena', None)
        self._nccl = False
        self.exchange = 'nccl'
        self.exchange_note = ''            # why the multicast exchange is not in use, when it was asked for
        self._mc = None
        if self.world > 1 and self.arena is not None:
            rt = self.arena.rt
            ident = None
            if self.rank == 0:
                buf = (np.zeros(128, dtype=np.uint8))
                rt.api.nccl_unique_id(buf.ctypes.data)
                ident = buf.tobytes()
            ident = self.comm.broadcast_bytes(ident, root=0)
            idbuf = np.frombuffer(ident, dtype=np.uint8).copy()
            rt.api.nccl_init(idbuf.ctypes.data, self.world, self.rank)
            self._nccl = True
        want = (exchange or os.environ.get('LG_DP_EXCHANGE', 'auto')).lower()
        if self.world == 1 and want == 'local' and self.arena is not None and hasattr(optimizer, '_bucket_step_range'):
            # one GPU, on request only: every bucket of parameters is updated by the small-footprint kernel on the
            # collective stream as soon as its gradients are final.  Measured on BERT-base batch 32 (profiles/
            # r2_bench_n1_local_vs_plain.txt): 8.97 ms per step against 8.53 ms for loss.backward(); optimizer.step() --
            # 3.7 GB of optimizer traffic running beside the GEMMs costs them more L2 / HBM than the overlap saves, so
            # 'auto' keeps the plain step on one GPU (with N ranks the traffic is 1/N and comes with the exchange)
            self.exchange = 'local'
        single = self.world == 1 and want == 'nvls'          # a team of one: exercises the whole path on one GPU
        if self.arena is not None and want in ('auto', 'nvls') and (self._nccl or single) \
                and hasattr(optimizer, '_mc_exchange_range'):
            if self._setup_multicast():
                self.exchange = 'nvls'
            elif want == 'nvls' and self.rank == 0:
                import sys
                sys.stderr.write("lightgrad_b200: NVLink multicast exchange unavailable (%s), using NCCL all-reduce\n"
                                 % self.exchange_note)
        if broadcast and self.world > 1:
            self.broadcast_parameters()

    def shard(self, n_rows):
        return shard_rows(n_rows, self.rank, self.world)

    def _all_ok(self, ok):
        """True only when every rank says so."""
        return self.comm.max_float(0.0 if ok else 1.0) == 0.0

    def _setup_multicast(self):
        """Create the multicast region (rank 0 creates the object, the others import it), bind this GPU's memory to
        it and move both optimizer arenas there.  Every rank takes the same decision; on any failure all fall back."""
        import ctypes as C
        a = self.arena
        api = a.rt.api
        yes = C.c_int(0)
        try:
            api.mc_supported(C.byref(yes))
        except Exception as exc:
            yes.value = 0
            self.exchange_note = 'lg_mc_supported: %s' % exc
        if not self._all_ok(bool(yes.value)):
            self.exchange_note = self.exchange_note or 'the device / driver reports no multicast support on some rank'
            return False
        region, goff, poff, foff = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
        created, err = False, None
        try:
            api.mc_region_bytes(a.total * 4, self.world, C.byref(region), C.byref(goff), C.byref(poff), C.byref(foff))
            fd = C.c_int(-1)
            if self.rank == 0:
                api.mc_create(region.value, self.world, C.byref(fd))
                created = True
        except Exception as exc:
            err = exc
        if not self._all_ok(err is None):
            self.exchange_note = 'creating the multicast object failed: %s' % err
            if created:
                api.mc_release()
            return False
        try:
            local_fd = _pass_fd(self.comm, fd.value)
            if self.rank != 0:
                api.mc_import(local_fd, region.value, self.world)
                created = True
            if local_fd >= 0:
                os.close(local_fd)
            api.mc_add_device()
        except Exception as exc:
            err = exc
        if not self._all_ok(err is None):          # (also the barrier cuMulticastBindMem needs: every device was added)
            self.exchange_note = 'sharing / joining the multicast object failed: %s' % err
            if created:
                api.mc_release()
            return False
        local, mcva = C.c_void_p(0), C.c_void_p(0)
        try:
            api.mc_bind(C.byref(local), C.byref(mcva))
        except Exception as exc:
            err = exc
        if not self._all_ok(err is None):
            self.exchange_note = 'binding memory to the multicast object failed: %s' % err
            api.mc_release()
            return False
        a.rebase(self.optimizer.parameters, local.value + poff.value, local.value + goff.value, keep=self)
        self._mc = (goff.value, poff.value, foff.value)
        self._buckets = None
        self.comm.barrier()
        return True

    def broadcast_parameters(self, root=0):
        if self._nccl:
            a = self.arena
            a.rt.api.nccl_broadcast(a.param_buf.ptr, a.total * 4, root)
            a.param_buf._bf16 = None
            return
        for p in self.optimizer.parameters:
            data = self.comm.broadcast_bytes(p.numpy().tobytes() if self.rank == root else None, root=root)
            new = np.frombuffer(data, dtype=p.dtype).reshape(p.shape)
            p.fill(0.0)
            p += p.__class__.from_numpy(new.copy(), requires_grad=False)

    def backward_and_step(self, loss, bucket_bytes=None):
        """``backward`` + ``optimizer.step()`` with the optimizer pipelined behind the exchange: the update of a
        bucket's parameters is queued on the communication stream right after that bucket's all-reduce, so only the
        last bucket's exchange and update remain after backward instead of the whole optimizer pass.  Safe because
        a bucket is exchanged only after every consumer of its parameters has run its backward."""
        # Measured (2 GPUs, local batch 128): 30.56 ms pipelined vs 30.44 ms plain -- whatever runs on the
        # communication stream competes with the persistent one-CTA-per-SM GEMMs of backward for the same SMs, so
        # moving the optimizer there buys nothing today.  Off unless LG_DP_PIPELINED_STEP=1.
        if self.exchange in ('nvls', 'local'):
            # nvls: gradient reduce-scatter + optimizer + parameter all-gather, one kernel per bucket, overlapped with
            # backward; local (one GPU): the optimizer alone, same kernel shape, same overlap
            self.backward(loss, bucket_bytes, _step_buckets=True)
            return
        fused = getattr(self.optimizer, '_fused_step_range', None)
        if self.world == 1 or not self._nccl or fused is None or os.environ.get('LG_DP_NO_OVERLAP') \
                or not os.environ.get('LG_DP_PIPELINED_STEP'):
            self.backward(loss, bucket_bytes)
            self.optimizer.step()
            return
        self.backward(loss, bucket_bytes, _step_buckets=True)

    def backward(self, loss, bucket_bytes=None, _step_buckets=False):
        """``loss.backward()`` with the gradient exchange overlapped: the flat gradient arena is cut into
        ~``bucket_bytes`` buckets of consecutive parameters; as soon as the walk has delivered the last
        contribution to every parameter of a bucket, its all-reduce is queued on the communication
        stream while the compute stream carries on with the rest of backward.  Backward reaches the
        parameters in reverse registration order, so buckets complete from the tail of the arena.
        The optimizer (compute stream) waits for the communication stream at the end."""
        if bucket_bytes is None:
            # measured at 8 GPUs (profiles/r2_bench_n8_nvls_b*.json): 24 / 48 / 96 MB buckets = 8.73 / 8.51 / 8.43 ms per
            # step -- every bucket is two cross-GPU barriers and a kernel beside the GEMMs of backward, and only the LAST
            # bucket (the word embeddings, final when backward ends) is exposed whatever the size
            bucket_bytes = int(os.environ.get('LG_DP_BUCKET_MB', '96' if self.exchange in ('nvls', 'local') else '64')) << 20
        nvls_step = _step_buckets and self.exchange == 'nvls'
        local_step = _step_buckets and self.exchange == 'local'
        if (nvls_step or local_step) and os.environ.get('LG_DP_NO_OVERLAP'):
            # the whole arena as ONE bucket after backward (A/B switch: how much does running beside backward buy?)
            loss.backward()
            a, P = self.arena, len(self.optimizer.parameters)
            a.adopt_grads(self.optimizer.parameters)
            a.rt.api.nccl_fork()
            if nvls_step:
                self.optimizer._mc_exchange_range(a, self._mc, 0, P, 0, a.total, self.rank, self.world, last=True)
            else:
                self.optimizer._bucket_step_range(a, 0, P, 0, a.total, last=True)
            a.rt.api.nccl_wait()
            a.param_buf._bf16 = None
            return
        if not (nvls_step or local_step) and (self.world == 1 or not self._nccl or os.environ.get('LG_DP_NO_OVERLAP')):
            loss.backward()
            self.sync_gradients()
            return
        from .autograd import Gradients
        a, params = self.arena, self.optimizer.parameters
        a.adopt_grads(params)
        if getattr(self, '_buckets', None) is None or getattr(self, '_bucket_bytes', None) != bucket_bytes:
            self._bucket_bytes = bucket_bytes
            # a parameter's slot runs to the next parameter's (64-element aligned) offset: the zero padding travels along
            ends = list(a.offsets[1:]) + [a.total]
            self._bucket_of, self._buckets = [], []      # per param -> bucket; bucket -> [lo, hi, n_params]
            # Backward finishes the arena from its END, so the buckets at its START are launched last and run exposed.
            # LG_DP_RAMP_BUCKETS=n closes the first n of them at a quarter of the size (the word embeddings alone, then
            # one encoder layer each).  Measured at 8 GPUs: 8.42 / 8.46 ms with n = 3 against 8.44 with equal buckets --
            # no difference, so the default keeps equal buckets.
            ramp = int(os.environ.get('LG_DP_RAMP_BUCKETS', '0')) if (nvls_step or local_step) else 0
            lo, count = 0, 0
            for i, hi in enumerate(ends):
                self._bucket_of.append(len(self._buckets))
                count += 1
                limit = bucket_bytes // 4 if len(self._buckets) < ramp else bucket_bytes
                if (hi - lo) * 4 >= limit or i == len(ends) - 1:
                    self._buckets.append([lo, hi, count])
                    lo, count = hi, 0
            self._index = {id(p): i for i, p in enumerate(params)}
        remaining = [b[2] for b in self._buckets]
        launched = [False] * len(self._buckets)
        api = a.rt.api

        n_buckets = len(self._buckets)
        first_param = [0] * n_buckets                      # bucket -> index of its first parameter
        for i, b in enumerate(self._bucket_of):
            if i == 0 or self._bucket_of[i - 1] != b:
                first_param[b] = i
        n_launched = [0]

        def launch(b):
            lo, hi, count = self._buckets[b]
            launched[b] = True
            n_launched[0] += 1
            api.nccl_fork()                                   # comm stream waits for the gradients written so far
            if nvls_step:
                self.optimizer._mc_exchange_range(a, self._mc, first_param[b], first_param[b] + count, lo, hi,
                                                  self.rank, self.world, last=n_launched[0] == n_buckets)
                return
            if local_step:
                self.optimizer._bucket_step_range(a, first_param[b], first_param[b] + count, lo, hi,
                                                  last=n_launched[0] == n_buckets)
                return
            api.nccl_allreduce_f32(a.grad_buf.ptr + lo * 4, hi - lo, 1, 1)
            if _step_buckets:
                # the optimizer update of this bucket, behind its all-reduce on the communication stream
                api.comm_compute_begin()
                try:
                    self.optimizer._fused_step_range(a, first_param[b], first_param[b] + count,
                                                     last=n_launched[0] == n_buckets)
                finally:
                    api.comm_compute_end()

        def leaf_done(t):
            i = self._index.get(id(t))
            if i is None:
                return
            b = self._bucket_of[i]
            remaining[b] -= 1
            if remaining[b] == 0 and not launched[b]:
                launch(b)
        prev = Gradients.leaf_hook
        Gradients.leaf_hook = leaf_done
        # optionally leave a few SMs to the collective (LG_DP_RESERVE_SMS): a persistent one-CTA-per-SM GEMM grid
        # that finds some SMs taken by NCCL's CTAs needs a second wave.  Measured at 8 GPUs: 9.72 / 9.85 / 9.83 /
        # 9.64 ms per step for 0 / 8 / 16 / 32 reserved SMs -- no gain, so the default reserves none.
        reserve = int(os.environ.get('LG_DP_RESERVE_SMS', '0'))
        if reserve > 0:
            api.gemm_sm_limit(max(a.rt.device_props()['sm_count'] - reserve, 1))
        try:
            loss.backward()
        finally:
            Gradients.leaf_hook = prev
            if reserve > 0:
                api.gemm_sm_limit(0)
        if self._mc is not None and os.environ.get('LG_MC_TRACE'):
            api.mc_trace_mark()                               # (measurement aid) the compute stream finished backward
        for b in range(len(self._buckets)):                   # parameters the loss does not reach
            if not launched[b]:
                launch(b)
        api.nccl_wait()                                       # compute stream waits for every bucket
        if _step_buckets:
            a.param_buf._bf16 = None                          # the parameters changed: a bf16 staging copy is stale

    def sync_gradients(self):
        """Average ``.grad`` of every parameter over the ranks (in place)."""
        if self.world == 1:
            return
        if self._nccl:
            a = self.arena
            a.adopt_grads(self.optimizer.parameters)
            a.rt.api.nccl_allreduce_f32(a.grad_buf.ptr, a.total, 1, 0)
            return
        for p in self.optimizer.parameters:
            g = p.grad
            avg = self.comm.allreduce_avg_numpy(g.numpy())
            g.fill(0.0)
            g += g.__class__.from_numpy(avg, requires_grad=False)

    def exchange_report(self, rt, comm, barrier, reps=5):
        """The gradient exchange on its own (timed with CUDA events, max over ranks), for bench.py's `comm` object."""
        a = self.arena
        if self.exchange == 'nvls':
            # the exchange kernel alone over the whole arena, without an update (kind 3: parameters written back as read)
            goff, poff, foff = self._mc
            nbytes = a.total * 4

            def run():
                rt.api.nccl_fork()
                rt.api.mc_exchange_step(3, goff, poff, foff, 0, a.total, self.rank, self.world, None, None, 0, None,
                                        None, 0.0, 0.0, 0.0, 0.0, 0.0, 0, 0)
                rt.api.nccl_wait()
            for _ in range(2):
                run()
            barrier()
            c0 = rt.Event().record()
            for _ in range(reps):
                run()
            c1 = rt.Event().record()
            c1.synchronize()
            ms = comm.max_float(c0.elapsed_ms(c1)) / reps
            return {'exchange': 'nvls', 'exchange_bytes_per_gpu_each_way': int(nbytes), 'exchange_ms_alone': round(ms, 3),
                    # per GPU the switch pulls the whole gradient arena (reduce-scatter) and pushes the whole
                    # parameter arena (all-gather): bytes / time is the NVLink rate per direction
                    'nvlink_gbps_per_direction': round(nbytes / (ms / 1e3) / 1e9, 1),
                    'note': 'multimem.ld_reduce reduce-scatter + multimem.st all-gather of the whole arena (no optimizer '
                            'update), alone on the GPU; in the step this work is fused with Adam and runs per bucket '
                            'next to the GEMMs of backward'}
        if not self._nccl:
            return {}
        nbytes = a.total * 4
        for _ in range(2):
            rt.api.nccl_allreduce_f32(a.grad_buf.ptr, a.total, 1, 0)
        barrier()
        c0 = rt.Event().record()
        for _ in range(reps):
            rt.api.nccl_allreduce_f32(a.grad_buf.ptr, a.total, 1, 0)
        c1 = rt.Event().record()
        c1.synchronize()
        ms = comm.max_float(c0.elapsed_ms(c1)) / reps
        return {'exchange': 'nccl', 'allreduce_bytes': int(nbytes), 'allreduce_ms_alone': round(ms, 3),
                'allreduce_busbw_gbps': round(2.0 * (self.world - 1) / self.world * nbytes / (ms / 1e3) / 1e9, 1),
                'note': 'ncclAllReduce of the whole gradient arena, alone on the GPU'}

    def close(self, timeout_s=20.0):
        """Tear the communicator down in order: every rank drains its streams, all ranks meet, then
        ncclCommDestroy.  Captured steps that hold NCCL nodes must have been destroyed before (StepGraph.destroy).
        ncclCommDestroy is called from a helper thread with a deadline: if a peer has already gone it can wait
        forever, and a clean interpreter exit matters more than returning NCCL's resources to a dying process."""
        if self._mc is not None:
            rt = self.arena.rt
            rt.synchronize()
            self.comm.barrier()              # nobody unbinds while a peer may still write through the switch
            self._mc = None
            self.exchange = 'nccl'
            rt.api.mc_release()
        if not self._nccl:
            return
        import threading
        self._nccl = False
        rt = self.arena.rt
        rt.synchronize()
        self.comm.barrier()
        done = []

        def destroy():
            try:
                rt.api.nccl_destroy()
                done.append(True)
            except Exception as exc:          # reported, not raised: the job's work is over
                done.append(exc)
        th = threading.Thread(target=destroy, daemon=True)
        th.start()
        th.join(timeout_s)
        if not done:
            import sys
            sys.stderr.write("lightgrad_b200: ncclCommDestroy did not return within %.0f s on rank %d; "
                             "leaving it to process exit\n" % (timeout_s, self.rank))
