"""Gradient mode switch and the reverse-mode graph walk.

Mirrors the public surface of the reference's ``Gradients`` class
(lightgrad/autograd/grads.py:4-42): ``no_grad()`` usable as context manager
and as decorator, ``disable()/enable()/_is_enabled()`` and
``backward(ctx, grad)``.

Deliberate divergence: the reference pops the last-inserted
context from an OrderedDict (grads.py:36), which is not a topological order --
a tensor with two consumers can be back-propagated before both consumers
contributed and then a second time, double counting.  Here every context is
visited exactly once, after all of its consumers, which is identical to the
reference on chains and trees (every graph the reference's own tests build).
"""
from functools import wraps


class _NoGrad(object):
    """Context manager / decorator that suspends graph recording."""
    __slots__ = ()

    def __enter__(self):
        Gradients._depth += 1

    def __exit__(self, *exc):
        d = Gradients._depth - 1
        Gradients._depth = d if d > 0 else 0

    def __call__(self, fn):
        @wraps(fn)
        def guarded(*args, **kwargs):
            Gradients._depth += 1
            try:
                return fn(*args, **kwargs)
            finally:
                d = Gradients._depth - 1
                Gradients._depth = d if d > 0 else 0
        return guarded


class Gradients(object):
    _depth = 0
    # keep ``.grad`` of intermediate (non-leaf) tensors after the walk, as the
    # reference does.  Trainers may switch it off to release memory early.
    retain_intermediate = True
    # optional callable(tensor): invoked by the outermost walk as soon as a leaf has received its last
    # gradient contribution (lets the data-parallel wrapper start reducing finished gradients while the
    # rest of backward is still running)
    leaf_hook = None
    # callables run when the OUTERMOST walk is over (a backend joins work it issued asynchronously, e.g. weight
    # gradients computed on a second stream)
    after_backward = []
    _walking = 0

    @staticmethod
    def disable():
        Gradients._depth += 1

    @staticmethod
    def enable():
        d = Gradients._depth - 1
        Gradients._depth = d if d > 0 else 0

    @staticmethod
    def _is_enabled() -> bool:
        return Gradients._depth == 0

    @staticmethod
    def no_grad():
        return _NoGrad()

    @staticmethod
    def _order(root):
        """Reverse post-order of the context DAG reachable from ``root``.

        Returns [(ctx, owner_tensor)], root first; owner_tensor is the tensor
        whose ``.ctx`` is that context (None for the root).  Iterative DFS so a
        12-layer BERT (~1k nodes deep) cannot hit the recursion limit.
        """
        post, seen = [], {id(root)}
        stack = [(root, None, iter(tuple(root.parent_tensors)))]
        while stack:
            ctx, owner, it = stack[-1]
            pushed = False
            for t in it:
                c = t.ctx
                if c is None or id(c) in seen:
                    continue
                seen.add(id(c))
                stack.append((c, t, iter(tuple(c.parent_tensors))))
                pushed = True
                break
            if not pushed:
                stack.pop()
                post.append((ctx, owner))
        post.reverse()
        return post

    @staticmethod
    def backward(ctx, grad, retain=None):
        """Propagate ``grad`` (gradient of the tensor produced by ``ctx``).

        ``retain=False`` releases the ``.grad`` of every non-leaf tensor once it
        has been consumed (used for the private graphs of WrapperFunctions so a
        second backward pass does not see stale partial sums).
        """
        order = Gradients._order(ctx)
        keep = Gradients.retain_intermediate if retain is None else retain
        hook = Gradients.leaf_hook if Gradients._walking == 0 else None
        pending = None
        wrappers_left, deferred = 0, []
        if hook is not None:
            # number of graph nodes that still owe each leaf a contribution
            pending = {}
            for node, _ in order:
                for t in node.parent_tensors:
                    if t.ctx is None:
                        pending[id(t)] = pending.get(id(t), 0) + 1
                # a WrapperFunction back-propagates through a private graph, which may reach leaves (parameters used by
                # closure) that this count cannot see: no leaf is reported as finished while such a node is outstanding
                if getattr(node, '_inner', None) is not None:
                    wrappers_left += 1
        Gradients._depth += 1
        Gradients._walking += 1
        try:
            for node, owner in order:
                g = grad if owner is None else owner.grad
                if g is None:
                    # nothing flowed into this node (e.g. a branch that only
                    # feeds non-differentiable consumers)
                    continue
                if not keep and owner is not None:
                    # this gradient dies right after the node is processed: hand its buffer to the
                    # node's backward, so a view of it (reshape / transpose backward) or the tensor
                    # itself (add backward) can be adopted by a parent instead of copied
                    owner._drop_grad()
                    g._temp = True
                node._backpropagate(g)
                if pending is not None:
                    if getattr(node, '_inner', None) is not None:
                        wrappers_left -= 1
                    for t in node.parent_tensors:
                        k = id(t)
                        if k in pending:
                            pending[k] -= 1
                            if pending[k] == 0:
                                del pending[k]
                                deferred.append(t)
                    if wrappers_left <= 0 and deferred:
                        for t in deferred:
                            hook(t)
                        deferred = []
            if pending is not None:
                for t in deferred:      # (a wrapper node that received no gradient was skipped above)
                    hook(t)
        finally:
            Gradients._walking -= 1
            d = Gradients._depth - 1
            Gradients._depth = d if d > 0 else 0
            if Gradients._walking == 0:
                for fn in Gradients.after_backward:
                    fn()
