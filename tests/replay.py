"""Replays the golden cases of tests/golden/*.npz (made by the real reference, see make_golden.py)
on any tensor class registered with lightgrad_b200's autograd core and yields (name, got, want)."""
import os
import ast
import numpy as np
import lightgrad_b200 as light
import lightgrad_b200.nn as nn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    with np.load(os.path.join(GOLD, name)) as z:
        return {k: z[k] for k in z.files}


def groups(g):
    out = {}
    for k, v in g.items():
        case, field = k.rsplit('/', 1)
        out.setdefault(case, {})[field] = v
    return out


SCALAR = {'x+2.5': lambda t: t + 2.5, '3-x': lambda t: 3 - t, 'x-1.5': lambda t: t - 1.5, 'x*0.3': lambda t: t * 0.3,
          'x_div_8': lambda t: t / 8.0, '2_div_x': lambda t: 2 / t, 'x**2': lambda t: t ** 2, 'x**0.5': lambda t: t ** 0.5,
          'x**-1': lambda t: t ** -1, 'x**1.7': lambda t: t ** 1.7}
REDUCE_KW = {'all': {}, 'ax0': dict(axis=0), 'ax1k': dict(axis=1, keepdims=True), 'ax2': dict(axis=2),
             'ax02': dict(axis=(0, 2)), 'ax12k': dict(axis=(1, 2), keepdims=True), 'ties': dict(axis=1)}
INDEX = {'transpose': lambda t: t.transpose(2, 0, 1), 'T': lambda t: t.transpose(),
         'reshape': lambda t: t.reshape(-1, 10), 'tr_reshape': lambda t: t.transpose(1, 0, 2).reshape(6, 20),
         'slice': lambda t: t[1:3, ::2, 1], 'int': lambda t: t[2], 'ellipsis': lambda t: t[..., 1:4],
         'gather': lambda t: t[np.array([3, 0, 3, 1])],
         'gather2': lambda t: t[np.array([0, 1, 3]), np.array([5, 0, 5])],
         'range_gather': lambda t: t[range(4), np.array([1, 0, 5, 5])], 'pad': lambda t: t.pad(2)}
DOT_T = {'2d': (False, False), 'ta': (True, False), 'tb': (False, True), 'batched': (False, False),
         'bcast': (False, False)}


def _grads(T, out, w, *leaves):
    for l in leaves:
        l.zero_grad()
    (out * T.from_numpy(w)).sum().backward()
    return [l.grad.numpy() for l in leaves]


def replay_ops(T, only=None):
    """Yields (case, field, got, want) for every op-level golden case."""
    G = groups(load('ops.npz'))
    for case, f in sorted(G.items()):
        if only is not None and not case.startswith(only):
            continue
        kind = case.split('/')
        if kind[0] == 'unary':
            x = T.from_numpy(f['x'])
            out = getattr(x, kind[1])()
            gx, = _grads(T, out, f['w'], x)
            yield case, 'out', out.numpy(), f['out']
            yield case, 'gx', gx, f['gx']
        elif kind[0] == 'binary':
            a, b = T.from_numpy(f['a']), T.from_numpy(f['b'])
            out = getattr(a, kind[1])(b)
            ga, gb = _grads(T, out, f['w'], a, b)
            yield case, 'out', out.numpy(), f['out']
            yield case, 'ga', ga, f['ga']
            yield case, 'gb', gb, f['gb']
        elif kind[0] == 'scalar':
            x = T.from_numpy(f['x'])
            out = SCALAR[kind[1]](x)
            gx, = _grads(T, out, f['w'], x)
            yield case, 'out', out.numpy(), f['out']
            yield case, 'gx', gx, f['gx']
        elif kind[0] == 'reduce':
            x = T.from_numpy(f['x'])
            out = getattr(x, kind[1])(**REDUCE_KW[kind[2]])
            gx, = _grads(T, out, f['w'], x)
            yield case, 'out', out.numpy(), f['out']
            yield case, 'gx', gx, f['gx']
        elif kind[0] == 'dot':
            ta, tb = DOT_T[kind[1]]
            a, b = T.from_numpy(f['a']), T.from_numpy(f['b'])
            out = (a.transpose(1, 0) if ta else a) @ (b.transpose(1, 0) if tb else b)
            ga, gb = _grads(T, out, f['w'], a, b)
            yield case, 'out', out.numpy(), f['out']
            yield case, 'ga', ga, f['ga']
            yield case, 'gb', gb, f['gb']
        elif kind[0] == 'index':
            x = T.from_numpy(f['x'])
            out = INDEX[kind[1]](x)
            gx, = _grads(T, out, f['w'], x)
            yield case, 'out', out.numpy(), f['out']
            yield case, 'gx', gx, f['gx']
        elif kind[0] == 'pool':
            x = T.from_numpy(f['x'])
            out = getattr(x, kind[1])()
            gx, = _grads(T, out, f['w'], x)
            yield case, 'out', out.numpy(), f['out']
            yield case, 'gx', gx, f['gx']
        elif kind[0] == 'conv':
            x, k = T.from_numpy(f['x']), T.from_numpy(f['k'])
            out = x.conv(k, strides=int(f['stride']))
            gx, gk = _grads(T, out, f['w'], x, k)
            yield case, 'out', out.numpy(), f['out']
            yield case, 'gx', gx, f['gx']
            yield case, 'gk', gk, f['gk']
        elif case == 'fused/softmax':
            x = T.from_numpy(f['x'])
            out = x.softmax(axis=-1)
            gx, = _grads(T, out, f['w'], x)
            yield case, 'out', out.numpy(), f['out']
            yield case, 'gx', gx, f['gx']
        elif case == 'fused/softmax_scaled':
            x = T.from_numpy(f['x'])
            if getattr(T, 'has_scaled_softmax', False):
                out = x.softmax(axis=-1, scale=float(f['scale']))
            else:
                out = (x / 4.0).softmax(axis=-1)
            gx, = _grads(T, out, f['w'], x)
            yield case, 'out', out.numpy(), f['out']
            yield case, 'gx', gx, f['gx']
        elif case == 'fused/layernorm':
            with nn.use_tensor(T):
                ln = nn.LayerNorm(16)
            ln.load_parameters({'weight': f['gamma'], 'bias': f['beta']})
            x = T.from_numpy(f['x'])
            out = ln(x)
            gx, gg, gb = _grads(T, out, f['w'], x, ln.weight, ln.bias)
            yield case, 'out', out.numpy(), f['out']
            yield case, 'gx', gx, f['gx']
            yield case, 'ggamma', gg, f['ggamma']
            yield case, 'gbeta', gb, f['gbeta']
        elif case == 'fused/gelu':
            from examples.bert import gelu
            x = T.from_numpy(f['x'])
            out = gelu(x)
            gx, = _grads(T, out, f['w'], x)
            yield case, 'out', out.numpy(), f['out']
            yield case, 'gx', gx, f['gx']
        elif case == 'loss/cross_entropy':
            x = T.from_numpy(f['logits'])
            loss = light.loss.cross_entropy(x, T.from_numpy(f['labels'], requires_grad=False))
            x.zero_grad()
            loss.backward()
            yield case, 'loss', loss.numpy(), f['loss']
            yield case, 'glogits', x.grad.numpy(), f['glogits']
        elif case == 'loss/mse':
            y = T.from_numpy(f['y'])
            loss = light.loss.mse(y, T.from_numpy(f['y_hat'], requires_grad=False))
            y.zero_grad()
            loss.backward()
            yield case, 'loss', loss.numpy(), f['loss']
            yield case, 'gy', y.grad.numpy(), f['gy']
        elif case == 'graph/diamond':
            x = T.from_numpy(f['x'])
            h = x.tanh()
            out = h.exp() + h
            x.zero_grad()
            out.sum().backward()
            yield case, 'out', out.numpy(), f['out']
            yield case, 'gx', x.grad.numpy(), f['gx']
        elif kind[0] == 'optim':
            mk = {'sgd': lambda ps, **kw: light.optim.SGD(ps, lr=0.1, **kw),
                  'sgd_momentum': lambda ps, **kw: light.optim.SGD(ps, lr=0.1, momentum=0.9, **kw),
                  'adam': lambda ps, **kw: light.optim.Adam(ps, lr=0.01, **kw),
                  'adabelief': lambda ps, **kw: light.optim.AdaBelief(ps, lr=0.01, **kw)}[kind[1]]
            for variant, kw in (('generic', dict(fused=False)), ('auto', {})):
                ps = [T.from_numpy(f['p0_0'].copy()), T.from_numpy(f['p0_1'].copy())]
                opt = mk(ps, **kw)
                if variant == 'auto' and opt.arena is None:
                    continue
                for s in range(5):
                    opt.zero_grad()
                    for i, p in enumerate(ps):
                        p.add_grad(T.from_numpy(f['g%d_%d' % (s, i)]))
                    opt.step()
                yield case, 'final_0/' + variant, ps[0].numpy(), f['final_0']
                yield case, 'final_1/' + variant, ps[1].numpy(), f['final_1']
        elif case == 'mnist':
            from examples import mnist as mn
            with nn.use_tensor(T):
                np.random.seed(0)
                model = mn.NN()
            model.load_parameters({'l1.weight': f['w1'], 'l2.weight': f['w2']})
            opt = light.optim.SGD(model.parameters(), lr=1e-4)
            x = T.from_numpy(f['x'], requires_grad=False)
            labels = T.from_numpy(f['labels'], requires_grad=False)
            losses = []
            for _ in range(5):
                losses.append(mn.train_step(model, opt, x, labels, T).item())
            yield case, 'losses', np.array(losses), f['losses']
            yield case, 'w1_final', model.l1.weight.numpy(), f['w1_final']
            yield case, 'w2_final', model.l2.weight.numpy(), f['w2_final']
        else:
            raise AssertionError("golden case %s has no replay rule" % case)


def replay_bert_tiny(T):
    """Forward + cross entropy + backward of the tiny BERT; yields (name, got, want)."""
    from examples import bert
    g = load('bert_tiny.npz')
    cfg = ast.literal_eval(str(g['cfg']))
    with nn.use_tensor(T):
        np.random.seed(0)
        model = bert.BertForMaskedLM(**cfg)
    model.load_parameters({k[len('param/'):]: v for k, v in g.items() if k.startswith('param/')})
    ids = T.from_numpy(g['ids'], requires_grad=False)
    labels = T.from_numpy(g['labels'], requires_grad=False)
    logits = model(ids)
    loss = light.loss.cross_entropy(logits.reshape(-1, cfg['vocab_size']), labels)
    for p in model.parameters():
        p.zero_grad()
    loss.backward()
    yield 'logits', logits.numpy(), g['logits']
    yield 'loss', loss.numpy(), g['loss']
    for n, p in model.named_parameters():
        yield 'grad/' + n, p.grad.numpy(), g['grad/' + n]
