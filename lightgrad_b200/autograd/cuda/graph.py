"""Whole-step CUDA graphs: capture one static-shape training step, replay it without python dispatch.

A BERT-base step issues ~1000 kernel launches from ~630 python-level operator calls; at ~20 us of
interpreter time per call the host, not the GPU, bounds the step ("hard parts").
The reference has nothing comparable (every OpenCL launch is followed by a blocking wait,
opencl/kernels.py:194).  Because shapes are static, the launch sequence of a step is identical every
iteration, so it can be recorded once and replayed as a single cudaGraphLaunch:

    step_graph = StepGraph(lambda: train_step(model, optimizer, ids, labels))   # runs + records once
    for _ in range(n):
        ids.copy_from_host(next_batch)        # refresh the static input tensors (outside the graph)
        loss = step_graph.replay()            # same tensor objects, new contents

Rules during capture: no host round trips (``numpy()``, ``item()``, ``from_numpy``) -- they raise.
State that changes per step must live on the device (Adam's step counter does).  Tensors created
inside the captured function live in an allocator pool private to the graph, so their addresses stay
reserved for replays; the tensors returned by the function are kept alive and refreshed by every replay.
"""
import ctypes as C
from . import runtime as rt


class StepGraph(object):

    def __init__(self, fn, warmup=2):
        """``fn()`` is run ``warmup`` times eagerly (lazy state, allocator warm-up), then once under capture."""
        self.fn = fn
        self._exec = None
        self._pool = C.c_int(0)
        for _ in range(warmup):
            fn()
        self.capture()

    def capture(self):
        api = rt.ensure_device()
        if self._exec is not None:
            api.graph_destroy(self._exec)
            self._exec = None
        n0 = rt.launch_count()
        from . import ops
        ops.drop_staging_copies()        # copies made before the capture would never be refreshed by replays
        api.graph_begin(C.byref(self._pool))
        try:
            self.outputs = self.fn()
        except BaseException:
            api.graph_abort()
            ops.drop_staging_copies()
            raise
        ops.drop_staging_copies()        # ... and copies made inside it live in the graph's private pool
        # drop the autograd history of the outputs: the activations it references go back to the graph's
        # private pool (their addresses stay reserved for replays) instead of staying allocated
        outs = self.outputs if isinstance(self.outputs, (tuple, list)) else (self.outputs,)
        for t in outs:
            if hasattr(t, 'detach'):
                t.detach()
        h, n = C.c_void_p(), C.c_uint64(0)
        api.graph_end(C.byref(h), C.byref(n))
        self._exec = h.value
        self.n_nodes = n.value
        self.n_kernels = rt.launch_count() - n0
        # the capture only recorded the step; run it once so the outputs hold real values
        api.graph_launch(self._exec, self.n_kernels)
        return self.outputs

    def replay(self):
        rt.api.graph_launch(self._exec, self.n_kernels)
        return self.outputs

    def destroy(self):
        """Release the instantiated graph now (teardown order matters when it holds NCCL nodes: graphs first,
        then the communicator).  The outputs keep their last values; ``replay`` is no longer possible."""
        if self._exec is not None and rt.api is not None:
            h, self._exec = self._exec, None
            rt.api.graph_destroy(h)

    def __del__(self):
        try:
            if self._exec is not None and rt.api is not None:
                rt.api.raw.lg_graph_destroy(self._exec)
        except Exception:
            pass
