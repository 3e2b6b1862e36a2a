"""Generates tests/golden/*.npz by running the REAL reference (ndoll1998/lightgrad, mounted read-only at
/root/reference) in the build container.  The GPU box has no /root/reference, so the vectors are
committed.  Re-run with:  python tests/golden/make_golden.py

How the reference is made importable and complete, all without touching its tree:
  * a stub `pyopencl` module is injected (the package imports pyopencl unconditionally);
  * three ops are re-registered on its CpuTensor through its own `register_op(..., overwrite=True)`:
    `sum` gains a backward, `dot` backward swaps the last two axes (not `.T`), `getitem` backward
    scatter-ADDs;
  * `Gradients.backward` is replaced by a reverse-topological walk (the stock LIFO walk double-counts
    shared nodes and cannot train BERT) -- identical on trees/chains;
  * bert.Embedding.forward becomes `self.weight[ids]` (stock code hard-codes `.cpu()[..].opencl()`).
Everything else -- every forward/backward formula, nn, loss, optim, the BERT and MNIST model code -- is
the reference's own code.
"""
import importlib.util
import os
import sys
import types
import numpy as np

REF = '/root/reference'
OUT = os.path.dirname(os.path.abspath(__file__))


def stub_pyopencl():
    cl = types.ModuleType('pyopencl')
    tools = types.ModuleType('pyopencl.tools')

    class _Enum(object):
        ACCELERATOR, CUSTOM, DEFAULT, ALL, CPU, GPU, TYPE = range(7)
    cl.device_type = _Enum
    cl.device_info = _Enum
    for n in ('Buffer', 'Device', 'Context', 'Kernel', 'CommandQueue', 'Program'):
        setattr(cl, n, type(n, (), {}))

    def get_platforms():
        raise RuntimeError("no OpenCL in the golden generator")
    cl.get_platforms = get_platforms
    for n in ('PooledBuffer', 'MemoryPool', 'ImmediateAllocator'):
        setattr(tools, n, type(n, (), {}))
    tools.dtype_to_ctype = lambda d: 'float'
    cl.tools = tools
    sys.modules['pyopencl'] = cl
    sys.modules['pyopencl.tools'] = tools


def load_reference():
    stub_pyopencl()
    sys.path.insert(0, REF)
    import lightgrad
    from lightgrad.autograd import CpuTensor, Gradients
    from lightgrad.autograd.func import Function
    from lightgrad.autograd.cpu.ops import _use_tensor_data

    @CpuTensor.register_op("sum", overwrite=True)
    @_use_tensor_data
    class sum_(Function):
        def forward(ctx, t, axis=None, keepdims=False):
            ctx.save_for_backward(t.shape, axis, keepdims)
            return t.sum(axis=axis, keepdims=keepdims)

        def backward(ctx, g):
            shape, axis, keepdims = ctx.get_saved_tensors()
            if axis is not None and not keepdims:
                g = np.expand_dims(g, axis=axis)
            return np.broadcast_to(g, shape).copy()
    sum_.__name__ = 'sum'

    @CpuTensor.register_op("dot", overwrite=True)
    @CpuTensor.register_op("__matmul__", overwrite=True)
    @_use_tensor_data
    class dot(Function):
        def forward(ctx, a, b):
            ctx.save_for_backward(a, b)
            return a @ b

        def backward(ctx, g):
            a, b = ctx.get_saved_tensors()
            return g @ np.swapaxes(b, -1, -2), np.swapaxes(a, -1, -2) @ g

    @CpuTensor.register_op("__getitem__", overwrite=True)
    @_use_tensor_data
    class getitem(Function):
        def forward(ctx, a, idx):
            if isinstance(idx, tuple):
                idx = tuple(t.data if isinstance(t, CpuTensor) else t for t in idx)
            ctx.save_for_backward(a.shape, idx)
            return a[idx]

        def backward(ctx, g):
            shape, idx = ctx.get_saved_tensors()
            out = np.zeros(shape, dtype=np.float32)
            np.add.at(out, idx, g)
            return out

    def topo_backward(ctx, grad):
        order, seen = [], {id(ctx)}
        stack = [(ctx, None, iter(list(ctx.parent_tensors)))]
        while stack:
            c, owner, it = stack[-1]
            for t in it:
                if t.ctx is not None and id(t.ctx) not in seen:
                    seen.add(id(t.ctx))
                    stack.append((t.ctx, t, iter(list(t.ctx.parent_tensors))))
                    break
            else:
                stack.pop()
                order.append((c, owner))
        for c, owner in reversed(order):
            g = grad if owner is None else owner.grad
            if g is None:
                continue
            with Gradients.no_grad():
                c._backpropagate(g)
    Gradients.backward = staticmethod(topo_backward)
    return lightgrad


def load_ref_example(name):
    spec = importlib.util.spec_from_file_location('ref_' + name, os.path.join(REF, 'examples', name + '.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    light = load_reference()
    from lightgrad.autograd import CpuTensor as T
    import lightgrad.nn as nn
    rs = np.random.RandomState(20261018)
    G = {}

    def rec(prefix, **arrays):
        for k, v in arrays.items():
            G[prefix + '/' + k] = np.asarray(v)

    def grads_of(out, w, *leaves):
        for l in leaves:
            l.zero_grad()
        (out * T.from_numpy(w)).sum().backward()
        return [l.grad.numpy().copy() for l in leaves]

    # ---- unary ops: forward values and gradients of sum(out * w)
    for name, lo, hi in (('neg', -2, 2), ('sin', -3, 3), ('cos', -3, 3), ('exp', -2, 2), ('log', 0.1, 5),
                         ('sigmoid', -4, 4), ('tanh', -3, 3), ('relu', -2, 2)):
        x = rs.uniform(lo, hi, size=(7, 13)).astype(np.float32)
        w = rs.uniform(-1, 1, size=(7, 13)).astype(np.float32)
        tx = T.from_numpy(x)
        out = getattr(tx, name)()
        gx, = grads_of(out, w, tx)
        rec('unary/' + name, x=x, w=w, out=out.numpy(), gx=gx)

    # ---- binary ops with broadcasting
    for name, lo, hi in (('add', -2, 2), ('sub', -2, 2), ('mul', -2, 2), ('div', 0.5, 3), ('pow', 0.5, 2)):
        for tag, sa, sb in (('same', (6, 10), (6, 10)), ('row', (6, 10), (1, 10)), ('col', (6, 10), (6, 1)),
                            ('lead', (3, 6, 10), (6, 10))):
            a = rs.uniform(lo, hi, size=sa).astype(np.float32)
            b = rs.uniform(lo, hi, size=sb).astype(np.float32)
            ta, tb = T.from_numpy(a), T.from_numpy(b)
            out = getattr(ta, name)(tb)
            w = rs.uniform(-1, 1, size=out.shape).astype(np.float32)
            ga, gb = grads_of(out, w, ta, tb)
            rec('binary/%s/%s' % (name, tag), a=a, b=b, w=w, out=out.numpy(), ga=ga, gb=gb)

    # ---- python-scalar operands through the operator sugar (generic wrappers on the CPU tensor)
    x = rs.uniform(0.5, 2, size=(5, 9)).astype(np.float32)
    w = rs.uniform(-1, 1, size=(5, 9)).astype(np.float32)
    for tag, fn in (('x+2.5', lambda t: t + 2.5), ('3-x', lambda t: 3 - t), ('x-1.5', lambda t: t - 1.5),
                    ('x*0.3', lambda t: t * 0.3), ('x_div_8', lambda t: t / 8.0), ('2_div_x', lambda t: 2 / t),
                    ('x**2', lambda t: t ** 2), ('x**0.5', lambda t: t ** 0.5), ('x**-1', lambda t: t ** -1),
                    ('x**1.7', lambda t: t ** 1.7)):
        tx = T.from_numpy(x)
        out = fn(tx)
        gx, = grads_of(out, w, tx)
        rec('scalar/' + tag, x=x, w=w, out=out.numpy(), gx=gx)

    # ---- reductions
    x = rs.uniform(-1, 1, size=(4, 6, 5)).astype(np.float32)
    for name in ('sum', 'max', 'min', 'mean'):
        for tag, kw in (('all', {}), ('ax0', dict(axis=0)), ('ax1k', dict(axis=1, keepdims=True)),
                        ('ax2', dict(axis=2)), ('ax02', dict(axis=(0, 2))), ('ax12k', dict(axis=(1, 2), keepdims=True))):
            tx = T.from_numpy(x)
            out = getattr(tx, name)(**kw)
            w = rs.uniform(-1, 1, size=out.shape).astype(np.float32)
            gx, = grads_of(out, w, tx)
            rec('reduce/%s/%s' % (name, tag), x=x, w=w, out=out.numpy(), gx=gx)
    # ties in max: every tied element receives the gradient (cpu/ops.py:268-272)
    xt = np.array([[1, 3, 3], [2, 2, 2]], dtype=np.float32)
    tx = T.from_numpy(xt)
    out = tx.max(axis=1)
    gx, = grads_of(out, np.array([1.0, 2.0], dtype=np.float32), tx)
    rec('reduce/max/ties', x=xt, w=np.array([1.0, 2.0], dtype=np.float32), out=out.numpy(), gx=gx)

    # ---- matmul: 2-D, transposed operands, batched, batch-broadcast (Linear on a 3-D input)
    for tag, sa, sb, ta_, tb_ in (('2d', (9, 14), (14, 11), False, False), ('ta', (14, 9), (14, 11), True, False),
                                  ('tb', (9, 14), (11, 14), False, True), ('batched', (3, 2, 9, 14), (3, 2, 14, 11), False, False),
                                  ('bcast', (3, 9, 14), (14, 11), False, False)):
        a = rs.uniform(-1, 1, size=sa).astype(np.float32)
        b = rs.uniform(-1, 1, size=sb).astype(np.float32)
        A, B = T.from_numpy(a), T.from_numpy(b)
        out = (A.transpose(1, 0) if ta_ else A) @ (B.transpose(1, 0) if tb_ else B)
        w = rs.uniform(-1, 1, size=out.shape).astype(np.float32)
        ga, gb = grads_of(out, w, A, B)
        rec('dot/' + tag, a=a, b=b, w=w, out=out.numpy(), ga=ga, gb=gb)

    # ---- transformations and indexing
    x = rs.uniform(-1, 1, size=(4, 6, 5)).astype(np.float32)
    for tag, fn in (('transpose', lambda t: t.transpose(2, 0, 1)), ('T', lambda t: t.transpose()),
                    ('reshape', lambda t: t.reshape(-1, 10)), ('tr_reshape', lambda t: t.transpose(1, 0, 2).reshape(6, 20)),
                    ('slice', lambda t: t[1:3, ::2, 1]), ('int', lambda t: t[2]), ('ellipsis', lambda t: t[..., 1:4]),
                    ('gather', lambda t: t[np.array([3, 0, 3, 1])]), ('gather2', lambda t: t[np.array([0, 1, 3]), np.array([5, 0, 5])]),
                    ('range_gather', lambda t: t[range(4), np.array([1, 0, 5, 5])]),
                    ('pad', lambda t: t.pad(2))):
        tx = T.from_numpy(x)
        out = fn(tx)
        w = rs.uniform(-1, 1, size=out.shape).astype(np.float32)
        gx, = grads_of(out, w, tx)
        rec('index/' + tag, x=x, w=w, out=out.numpy(), gx=gx)
    # pooling: 4-D input (the reference's pool backward permutation is only right when ndim == 2*len(kernel))
    x4 = rs.uniform(-1, 1, size=(2, 3, 6, 5)).astype(np.float32)
    for tag, fn in (('max_pool', lambda t: t.max_pool()), ('min_pool', lambda t: t.min_pool()),
                    ('mean_pool', lambda t: t.mean_pool())):
        tx = T.from_numpy(x4)
        out = fn(tx)
        w = rs.uniform(-1, 1, size=out.shape).astype(np.float32)
        gx, = grads_of(out, w, tx)
        rec('pool/' + tag, x=x4, w=w, out=out.numpy(), gx=gx)
    # convolution (cpu/ops.py:298-356)
    xc = rs.uniform(-1, 1, size=(3, 2, 7, 6)).astype(np.float32)
    kc = rs.uniform(-1, 1, size=(4, 2, 3, 3)).astype(np.float32)
    for tag, st in (('s1', 1), ('s2', 2)):
        tx, tk = T.from_numpy(xc), T.from_numpy(kc)
        out = tx.conv(tk, strides=st)
        w = rs.uniform(-1, 1, size=out.shape).astype(np.float32)
        gx, gk = grads_of(out, w, tx, tk)
        rec('conv/' + tag, x=xc, k=kc, w=w, out=out.numpy(), gx=gx, gk=gk, stride=np.int64(st))

    # ---- softmax / LayerNorm / gelu / losses
    x = rs.uniform(-2, 2, size=(3, 5, 16)).astype(np.float32)
    tx = T.from_numpy(x)
    out = tx.softmax(axis=-1)
    w = rs.uniform(-1, 1, size=out.shape).astype(np.float32)
    gx, = grads_of(out, w, tx)
    rec('fused/softmax', x=x, w=w, out=out.numpy(), gx=gx)
    tx = T.from_numpy(x)
    out = (tx / 4.0).softmax(axis=-1)
    gx, = grads_of(out, w, tx)
    rec('fused/softmax_scaled', x=x, w=w, out=out.numpy(), gx=gx, scale=np.float32(0.25))
    ln = nn.LayerNorm(16)
    gam = rs.uniform(0.5, 1.5, size=(16,)).astype(np.float32)
    bet = rs.uniform(-0.5, 0.5, size=(16,)).astype(np.float32)
    ln.load_parameters({'weight': gam, 'bias': bet})
    tx = T.from_numpy(x)
    out = ln(tx)
    gx, gg, gb = grads_of(out, w, tx, ln.weight, ln.bias)
    rec('fused/layernorm', x=x, w=w, gamma=gam, beta=bet, out=out.numpy(), gx=gx, ggamma=gg, gbeta=gb)
    bert = load_ref_example('bert')
    tx = T.from_numpy(x)
    out = bert.gelu(tx)
    gx, = grads_of(out, w, tx)
    rec('fused/gelu', x=x, w=w, out=out.numpy(), gx=gx)
    logits = rs.uniform(-3, 3, size=(12, 37)).astype(np.float32)
    labels = rs.randint(0, 37, size=(12,)).astype(np.int32)
    tl = T.from_numpy(logits)
    loss = light.loss.cross_entropy(tl, T.from_numpy(labels, requires_grad=False))
    tl.zero_grad()
    loss.backward()
    rec('loss/cross_entropy', logits=logits, labels=labels, loss=loss.numpy(), glogits=tl.grad.numpy())
    y = rs.uniform(-1, 1, size=(8, 10)).astype(np.float32)
    yh = rs.uniform(-1, 1, size=(8, 10)).astype(np.float32)
    ty = T.from_numpy(y)
    loss = light.loss.mse(ty, T.from_numpy(yh, requires_grad=False))
    ty.zero_grad()
    loss.backward()
    rec('loss/mse', y=y, y_hat=yh, loss=loss.numpy(), gy=ty.grad.numpy())

    # ---- a diamond graph (shared node): pins the topological walk
    x = rs.uniform(-1, 1, size=(4, 4)).astype(np.float32)
    tx = T.from_numpy(x)
    h = tx.tanh()
    out = h.exp() + h
    tx.zero_grad()
    out.sum().backward()
    rec('graph/diamond', x=x, out=out.numpy(), gx=tx.grad.numpy())

    # ---- optimizers: 5 steps on two parameters with fixed gradients sequence
    p0 = [rs.uniform(-1, 1, size=(5, 3)).astype(np.float32), rs.uniform(-1, 1, size=(7,)).astype(np.float32)]
    gseq = [[rs.uniform(-1, 1, size=p.shape).astype(np.float32) for p in p0] for _ in range(5)]
    for tag, make in (('sgd', lambda ps: light.optim.SGD(ps, lr=0.1)),
                      ('sgd_momentum', lambda ps: light.optim.SGD(ps, lr=0.1, momentum=0.9)),
                      ('adam', lambda ps: light.optim.Adam(ps, lr=0.01)),
                      ('adabelief', lambda ps: light.optim.AdaBelief(ps, lr=0.01))):
        ps = [T.from_numpy(p.copy()) for p in p0]
        opt = make(ps)
        for gs in gseq:
            opt.zero_grad()
            for p, g in zip(ps, gs):
                p.add_grad(T.from_numpy(g))
            opt.step()
        rec('optim/' + tag, p0_0=p0[0], p0_1=p0[1], final_0=ps[0].numpy(), final_1=ps[1].numpy(),
            **{'g%d_%d' % (s, i): g for s, gs in enumerate(gseq) for i, g in enumerate(gs)})

    # ---- MNIST MLP (config 1): 5 steps of the reference loop, mse + SGD(lr=1e-4)
    mn = load_ref_example_mnist()
    np.random.seed(0)
    model = mn.NN()
    w1, w2 = model.l1.weight.numpy().copy(), model.l2.weight.numpy().copy()
    rs2 = np.random.RandomState(0)
    xb = rs2.uniform(0, 1, size=(64, 1, 28, 28)).astype(np.float32)
    yb = rs2.randint(0, 10, size=(64,)).astype(np.int16)
    opt = light.optim.SGD(model.parameters(), lr=1e-4)
    losses = []
    for _ in range(5):
        y = model(T.from_numpy(xb, requires_grad=False))
        one_hot = light.zeros((64, 10))
        one_hot[range(64), T.from_numpy(yb, requires_grad=False)] = 1
        l = light.loss.mse(y, one_hot)
        opt.zero_grad()
        l.backward()
        opt.step()
        losses.append(l.item())
    rec('mnist', w1=w1, w2=w2, x=xb, labels=yb, losses=np.array(losses, dtype=np.float64),
        w1_final=model.l1.weight.numpy(), w2_final=model.l2.weight.numpy())

    # ---- tiny BERT (2 layers, H=32, 4 heads, I=64, V=100, B=2, S=8): loss + every parameter gradient
    bert.Embedding.forward = lambda self, ids: self.weight[ids]
    cfg = dict(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4, vocab_size=100,
               max_position_embeddings=16, type_vocab_size=2, attention_probs_dropout_prob=0.0, hidden_dropout_prob=0.0)
    np.random.seed(0)
    model = bert.BertForMaskedLM(**cfg)
    # xavier on these tiny shapes gives near-zero activations; scale up so the test is sensitive
    params = {}
    for n, p in model.named_parameters():
        a = p.numpy().copy()
        if 'LayerNorm' not in n and not n.endswith('predictions.bias'):
            a = (a * 8).astype(np.float32)
        params[n] = a
    model.load_parameters(params)
    ids = np.random.RandomState(1).randint(0, 100, size=(2, 8)).astype(np.int32)
    ids[0, 3] = ids[0, 5]          # a repeated token exercises scatter-add
    labels = np.random.RandomState(2).randint(0, 100, size=(16,)).astype(np.int32)
    logits = model(T.from_numpy(ids, requires_grad=False))
    loss = light.loss.cross_entropy(logits.reshape(-1, 100), T.from_numpy(labels, requires_grad=False))
    for p in model.parameters():
        p.zero_grad()
    loss.backward()
    out = {'cfg': np.array(repr(cfg)), 'ids': ids, 'labels': labels, 'loss': loss.numpy(), 'logits': logits.numpy()}
    for n, p in model.named_parameters():
        out['param/' + n] = params[n]
        out['grad/' + n] = p.grad.numpy()
    np.savez_compressed(os.path.join(OUT, 'bert_tiny.npz'), **out)

    np.savez_compressed(os.path.join(OUT, 'ops.npz'), **G)
    print("wrote", len(G), "op vectors and", len(out), "bert vectors to", OUT)


def load_ref_example_mnist():
    # examples/mnist.py imports matplotlib and tqdm at module level; neither is needed for the model
    for name in ('matplotlib', 'matplotlib.pyplot', 'tqdm'):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.trange = range
            sys.modules[name] = m
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    return load_ref_example('mnist')


if __name__ == '__main__':
    main()
