"""Grouped matmul (lg_gemm_grouped) and the fused self-attention node against the oracle.

The node replaces BertSelfAttention.forward of the reference (examples/bert.py:60-93): the oracle side
below spells that forward with the reference's own operators on the CPU tensor, so output and every
gradient are compared op-for-op.  Exact mode: <= 1e-5 (fp32 matmul bound); tf32 mode: <= 5e-3.
"""
import math
import numpy as np
import pytest
from lightgrad_b200 import CudaTensor
from lightgrad_b200.autograd.cuda import ops, runtime as rt
from lightgrad_b200.autograd.cuda.ops import _gemm_grouped, _swap_last
from oracle import CpuTensor


@pytest.fixture(params=["fake", pytest.param("gpu", marks=pytest.mark.gpu)])
def device(request):
    request.getfixturevalue("fake_device" if request.param == "fake" else "cuda")
    return request.param


@pytest.fixture(params=["fp32", "tf32"])
def mode(request, device):
    prev = ops.set_matmul_mode(request.param)
    yield request.param
    ops.set_matmul_mode(prev)


def rel(got, want):
    return float(np.abs(got - want).max() / (np.abs(want).max() + 1e-30))


def _tol(mode):
    return 1e-5 if mode == 'fp32' else 5e-3


@pytest.mark.parametrize("shape", [(256, 192, 128), (512, 768, 256), (100, 132, 72)])
def test_grouped_independent_results(mode, shape):
    M, N, K = shape
    rs = np.random.RandomState(3)
    a = [rs.uniform(-1, 1, (M, K)).astype(np.float32) for _ in range(3)]
    w = [rs.uniform(-1, 1, (N, K)).astype(np.float32) for _ in range(3)]        # used transposed (MN-major B)
    bias = [rs.uniform(-1, 1, (N,)).astype(np.float32) for _ in range(3)]
    A, W, B = ([CudaTensor.from_numpy(v) for v in vs] for vs in (a, w, bias))
    out = CudaTensor.empty((3, M, N))
    parts = [out._view((M, N), (N, 1), g * M * N) for g in range(3)]
    _gemm_grouped(A, [_swap_last(t) for t in W], parts, B)
    got = out.numpy()
    for g in range(3):
        want = a[g].astype(np.float64) @ w[g].T.astype(np.float64) + bias[g]
        assert rel(got[g], want) <= _tol(mode), (shape, g)
    # accumulate: every group adds into what its result already holds
    _gemm_grouped(A, [_swap_last(t) for t in W], parts, accumulate=True)
    got2 = out.numpy()
    for g in range(3):
        want = 2 * (a[g].astype(np.float64) @ w[g].T.astype(np.float64)) + bias[g]
        assert rel(got2[g], want) <= _tol(mode), (shape, g, 'acc')


@pytest.mark.parametrize("shape", [(256, 192, 128), (384, 768, 768), (100, 132, 72)])
@pytest.mark.parametrize("groups", [2, 3, 4])
def test_grouped_shared_result_is_k_concatenation(mode, shape, groups):
    M, N, K = shape
    rs = np.random.RandomState(4)
    a = [rs.uniform(-1, 1, (M, K)).astype(np.float32) for _ in range(groups)]
    b = [rs.uniform(-1, 1, (K, N)).astype(np.float32) for _ in range(groups)]
    A, B = ([CudaTensor.from_numpy(v) for v in vs] for vs in (a, b))
    out = CudaTensor.empty((M, N))
    _gemm_grouped(A, B, [out] * groups)
    want = sum(x.astype(np.float64) @ y.astype(np.float64) for x, y in zip(a, b))
    assert rel(out.numpy(), want) <= _tol(mode)
    _gemm_grouped(A, B, [out] * groups, accumulate=True)
    assert rel(out.numpy(), 2 * want) <= _tol(mode)


def test_grouped_rejects_partially_shared_results(device):
    if device == 'fake':
        pytest.skip("argument check lives in the CUDA library")
    prev = ops.set_matmul_mode('tf32')
    try:
        A = [CudaTensor.zeros((128, 128)) for _ in range(3)]
        o1, o2 = CudaTensor.zeros((128, 128)), CudaTensor.zeros((128, 128))
        with pytest.raises(RuntimeError):
            _gemm_grouped(A, A, [o1, o1, o2])
    finally:
        ops.set_matmul_mode(prev)


def _oracle_attention(T, x, wq, bq, wk, bk, wv, bv, heads):
    # BertSelfAttention.forward with the reference's operator set (examples/bert.py:60-93)
    b, s, H = x.shape
    d = H // heads
    Q = x @ wq.transpose(1, 0) + bq
    K = x @ wk.transpose(1, 0) + bk
    V = x @ wv.transpose(1, 0) + bv
    Q = Q.reshape(b, s, heads, d).transpose(0, 2, 1, 3)
    K = K.reshape(b, s, heads, d).transpose(0, 2, 3, 1)
    V = V.reshape(b, s, heads, d).transpose(0, 2, 1, 3)
    P = (Q @ K / math.sqrt(d)).softmax(axis=-1)
    return (P @ V).transpose(0, 2, 1, 3).reshape(b, s, H)


# (4,100,128,2): ragged rows inside the fused softmax epilogue; (4,64,128,2): its 64-column tile; (2,132,128,2): too
# long for the row epilogue -> separate softmax kernels
# (2,128,128,2) and (3,128,768,12): seq 128 x head_dim 64 -> the one-kernel attention (lg_attention_fwd / _bwd) in tf32 mode
@pytest.mark.parametrize("cfg", [(2, 16, 64, 4), (3, 128, 256, 4), (2, 24, 96, 3), (4, 100, 128, 2), (4, 64, 128, 2),
                                 (2, 132, 128, 2), (2, 128, 128, 2), (3, 128, 768, 12)])
@pytest.mark.parametrize("preset_grads", [False, True])
def test_self_attention_matches_oracle(mode, cfg, preset_grads):
    b, s, H, heads = cfg
    rs = np.random.RandomState(5)
    x = rs.uniform(-1, 1, (b, s, H)).astype(np.float32)
    ws = [(rs.uniform(-1, 1, (H, H)) / math.sqrt(H)).astype(np.float32) for _ in range(3)]
    bs = [rs.uniform(-0.1, 0.1, (H,)).astype(np.float32) for _ in range(3)]
    up = rs.uniform(-1, 1, (b, s, H)).astype(np.float32)

    def run(T):
        X = T.from_numpy(x)
        W = [T.from_numpy(w) for w in ws]
        B = [T.from_numpy(v) for v in bs]
        if T is CudaTensor:
            if preset_grads:
                # gradients already allocated (optimizer arena): the node adds into them in place
                for p in W + B:
                    p.zero_grad()
            out = X.self_attention(W[0], B[0], W[1], B[1], W[2], B[2], heads=heads)
        else:
            out = _oracle_attention(T, X, W[0], B[0], W[1], B[1], W[2], B[2], heads)
        (out * T.from_numpy(up, requires_grad=False)).sum().backward()
        return out.numpy(), [X.grad.numpy()] + [p.grad.numpy() for p in W + B]

    want_out, want_g = run(CpuTensor)
    got_out, got_g = run(CudaTensor)
    tol = _tol(mode) * (10 if mode == 'fp32' else 1)       # composite of ~8 fp32 matmuls + softmax
    assert rel(got_out, want_out) <= tol
    names = ['x', 'wq', 'wk', 'wv', 'bq', 'bk', 'bv']
    gmax = max(float(np.abs(g).max()) for g in want_g)
    for n, g, w in zip(names, got_g, want_g):
        assert g.shape == w.shape, n
        # key.bias has a mathematically zero gradient (softmax is shift invariant): compare on the common scale
        assert float(np.abs(g - w).max()) <= tol * max(float(np.abs(w).max()), 1e-3 * gmax), n


def test_bert_layer_uses_fused_node_and_matches_unfused(device):
    import lightgrad_b200.nn as nn
    from examples.bert import BertSelfAttention
    with nn.use_tensor(CudaTensor):
        np.random.seed(7)
        att = BertSelfAttention(64, 4)
    x = np.random.uniform(-1, 1, (2, 16, 64)).astype(np.float32)
    X1, X2 = CudaTensor.from_numpy(x), CudaTensor.from_numpy(x)
    fused, none = att(X1, need_probs=False)
    assert none is None and fused.ctx.__class__.__name__ == 'self_attention'
    plain, probs = att(X2)
    assert probs is not None and probs.shape == (2, 4, 16, 16)
    np.testing.assert_allclose(fused.numpy(), plain.numpy(), rtol=1e-5, atol=1e-6)
    fused.sum().backward()
    g1 = {k: p.grad.numpy().copy() for k, p in att.named_parameters()}
    for _, p in att.named_parameters():
        p.zero_grad()
    plain.sum().backward()
    for k, p in att.named_parameters():
        np.testing.assert_allclose(g1[k], p.grad.numpy(), rtol=1e-4, atol=1e-5, err_msg=k)
    np.testing.assert_allclose(X1.grad.numpy(), X2.grad.numpy(), rtol=1e-4, atol=1e-5)


def _oracle_mlp(T, x, w1, b1, w2, b2):
    # BertLayer's feed-forward with the reference's operators (examples/bert.py:12,150-153)
    h = x @ w1.transpose(1, 0) + b1
    a = 0.5 * h * (1.0 + (h * 0.7978845608 * (1.0 + 0.044715 * h * h)).tanh())
    return a @ w2.transpose(1, 0) + b2


@pytest.mark.parametrize("cfg", [(2, 16, 64, 256), (4, 128, 256, 1024), (3, 20, 96, 200)])
@pytest.mark.parametrize("preset_grads", [False, True])
def test_mlp_gelu_matches_oracle(mode, cfg, preset_grads):
    b, s, H, F = cfg
    rs = np.random.RandomState(6)
    x = rs.uniform(-1, 1, (b, s, H)).astype(np.float32)
    w1 = (rs.uniform(-1, 1, (F, H)) * 2 / math.sqrt(H)).astype(np.float32)
    w2 = (rs.uniform(-1, 1, (H, F)) * 2 / math.sqrt(F)).astype(np.float32)
    b1 = rs.uniform(-0.5, 0.5, (F,)).astype(np.float32)
    b2 = rs.uniform(-0.5, 0.5, (H,)).astype(np.float32)
    up = rs.uniform(-1, 1, (b, s, H)).astype(np.float32)

    def run(T):
        X = T.from_numpy(x)
        P = [T.from_numpy(v) for v in (w1, b1, w2, b2)]
        if T is CudaTensor:
            if preset_grads:
                for p in P:
                    p.zero_grad()
                X.zero_grad()               # dX is then reduce-added into the existing gradient
            out = X.mlp_gelu(*P)
        else:
            out = _oracle_mlp(T, X, *P)
        (out * T.from_numpy(up, requires_grad=False)).sum().backward()
        return out.numpy(), [X.grad.numpy()] + [p.grad.numpy() for p in P]

    want_out, want_g = run(CpuTensor)
    got_out, got_g = run(CudaTensor)
    tol = _tol(mode) * (10 if mode == 'fp32' else 1)
    assert rel(got_out, want_out) <= tol
    for n, g, w in zip(['x', 'w1', 'b1', 'w2', 'b2'], got_g, want_g):
        assert g.shape == w.shape, n
        assert rel(g, w) <= tol, n


@pytest.mark.parametrize("shape", [(4, 16, 64), (2, 128, 768), (3, 7, 100)])
@pytest.mark.parametrize("preset_grads", [False, True])
def test_add_layernorm_matches_oracle(device, shape, preset_grads):
    # LayerNorm(hidden + residual) of BertAttention / BertLayer (examples/bert.py:113,158) with the reference's
    # own composition on the CPU tensor as the oracle (nn.py:109-124)
    rs = np.random.RandomState(8)
    a = rs.uniform(-1, 1, shape).astype(np.float32)
    b = rs.uniform(-1, 1, shape).astype(np.float32)
    w = rs.uniform(0.5, 1.5, shape[-1:]).astype(np.float32)
    bias = rs.uniform(-0.5, 0.5, shape[-1:]).astype(np.float32)
    up = rs.uniform(-1, 1, shape).astype(np.float32)

    def run(T):
        A, B, W, Bi = (T.from_numpy(v) for v in (a, b, w, bias))
        if T is CudaTensor:
            if preset_grads:
                W.zero_grad()
                Bi.zero_grad()
            out = A.add_layernorm(B, W, Bi, eps=1e-5)
        else:
            x = A + B
            D = x - x.mean(axis=-1, keepdims=True)
            V = (D * D).mean(axis=-1, keepdims=True)
            out = D / (V + 1e-5).pow(1 / 2) * W + Bi
        (out * T.from_numpy(up, requires_grad=False)).sum().backward()
        return out.numpy(), [t.grad.numpy() for t in (A, B, W, Bi)]

    want_out, want_g = run(CpuTensor)
    got_out, got_g = run(CudaTensor)
    np.testing.assert_allclose(got_out, want_out, rtol=3e-5, atol=3e-6)
    for n, g, wv in zip('abwB', got_g, want_g):
        np.testing.assert_allclose(g, wv, rtol=2e-4, atol=2e-5 * max(1.0, float(np.abs(wv).max())), err_msg=n)


@pytest.mark.gpu
def test_side_stream_results_visible_after_join(cuda):
    # work issued between lg_side_begin / lg_side_end runs on the second stream; lg_side_join (and, implicitly,
    # numpy()) orders the compute stream after it
    prev = ops.set_matmul_mode('tf32')
    try:
        rs = np.random.RandomState(11)
        a = rs.uniform(-1, 1, (512, 256)).astype(np.float32)
        b = rs.uniform(-1, 1, (256, 384)).astype(np.float32)
        A, B = CudaTensor.from_numpy(a), CudaTensor.from_numpy(b)
        acc = CudaTensor.zeros((512, 384))
        for _ in range(3):
            with rt.side_stream(A, B, writes=(acc,)):
                ops._gemm(A, B, out=acc, accumulate=True)
            # an in-place update from the compute stream must first wait for the side stream's writes
            acc += 1.0
        rt.side_join()
        want = 3 * (a.astype(np.float64) @ b.astype(np.float64)) + 3
        assert rel(acc.numpy(), want) <= 5e-3
    finally:
        ops.set_matmul_mode(prev)


@pytest.mark.gpu
def test_gemm_sm_limit_keeps_results(cuda):
    prev = ops.set_matmul_mode('tf32')
    try:
        rs = np.random.RandomState(12)
        a = rs.uniform(-1, 1, (1024, 512)).astype(np.float32)
        b = rs.uniform(-1, 1, (512, 768)).astype(np.float32)
        A, B = CudaTensor.from_numpy(a), CudaTensor.from_numpy(b)
        want = a.astype(np.float64) @ b.astype(np.float64)
        for limit in (100, 37, 1, 0):
            rt.api.gemm_sm_limit(limit)
            assert rel((A @ B).numpy(), want) <= 5e-3, limit
    finally:
        rt.api.gemm_sm_limit(0)
        ops.set_matmul_mode(prev)


def test_mlp_gelu_unaligned_width_takes_the_exact_path(mode):
    # intermediate width not a multiple of 4: no TMA-aligned pitch -> product and activation as separate kernels
    rs = np.random.RandomState(13)
    x = rs.uniform(-1, 1, (2, 8, 64)).astype(np.float32)
    w1 = (rs.uniform(-1, 1, (130, 64)) / 8).astype(np.float32)
    w2 = (rs.uniform(-1, 1, (64, 130)) / 11).astype(np.float32)
    b1 = rs.uniform(-0.5, 0.5, (130,)).astype(np.float32)
    b2 = rs.uniform(-0.5, 0.5, (64,)).astype(np.float32)

    def run(T):
        X = T.from_numpy(x)
        P = [T.from_numpy(v) for v in (w1, b1, w2, b2)]
        out = X.mlp_gelu(*P) if T is CudaTensor else _oracle_mlp(T, X, *P)
        out.sum().backward()
        return out.numpy(), [X.grad.numpy()] + [p.grad.numpy() for p in P]

    want_out, want_g = run(CpuTensor)
    got_out, got_g = run(CudaTensor)
    tol = _tol(mode) * (10 if mode == 'fp32' else 1)
    assert rel(got_out, want_out) <= tol
    for g, w in zip(got_g, want_g):
        assert rel(g, w) <= tol


@pytest.mark.parametrize("preset_grads", [False, True])
def test_tied_embedding_and_projection_weight(mode, preset_grads):
    # one table used as an embedding (gather / scatter-add backward, compute stream) AND as a Linear weight (dW on the
    # side stream): both contributions must land in the same gradient
    rs = np.random.RandomState(14)
    V, H, n = 96, 64, 256
    w = (rs.uniform(-1, 1, (V, H)) / 8).astype(np.float32)
    ids = rs.randint(0, V, size=(n,)).astype(np.int32)
    up = rs.uniform(-1, 1, (n, V)).astype(np.float32)

    def run(T):
        W = T.from_numpy(w)
        if T is CudaTensor and preset_grads:
            W.zero_grad()
        x = W[ids]
        y = x.linear(W) if T is CudaTensor else x @ W.transpose(1, 0)
        (y * T.from_numpy(up, requires_grad=False)).sum().backward()
        return y.numpy(), W.grad.numpy()

    want_y, want_g = run(CpuTensor)
    got_y, got_g = run(CudaTensor)
    tol = _tol(mode) * (10 if mode == 'fp32' else 1)
    assert rel(got_y, want_y) <= tol
    assert rel(got_g, want_g) <= tol


@pytest.mark.gpu
def test_fused_attention_kernel_equals_the_composed_path_at_bert_batch(cuda):
    """384 (batch, head) tiles -- more than the persistent grids hold at once, so every CTA loops over several tiles --
    against the batched-GEMM path with the same tf32 products (lg_gemm + softmax epilogues)."""
    b, s, H, heads = 32, 128, 768, 12
    rs = np.random.RandomState(11)
    x = rs.uniform(-1, 1, (b, s, H)).astype(np.float32)
    ws = [(rs.uniform(-1, 1, (H, H)) / math.sqrt(H)).astype(np.float32) for _ in range(3)]
    bs = [rs.uniform(-0.1, 0.1, (H,)).astype(np.float32) for _ in range(3)]
    up = rs.uniform(-1, 1, (b, s, H)).astype(np.float32)
    prev = ops.set_matmul_mode('tf32')
    try:
        def run(fused):
            (ops._DISABLED.discard if fused else ops._DISABLED.add)('attn_fused')
            X = CudaTensor.from_numpy(x)
            W = [CudaTensor.from_numpy(w) for w in ws]
            B = [CudaTensor.from_numpy(v) for v in bs]
            n0 = rt.launch_count()
            out = X.self_attention(W[0], B[0], W[1], B[1], W[2], B[2], heads=heads)
            (out * CudaTensor.from_numpy(up, requires_grad=False)).sum().backward()
            return out.numpy(), [X.grad.numpy()] + [p.grad.numpy() for p in W + B], rt.launch_count() - n0
        want_out, want_g, n_composed = run(False)
        got_out, got_g, n_fused = run(True)
    finally:
        ops._DISABLED.discard('attn_fused')
        ops.set_matmul_mode(prev)
    assert n_fused < n_composed                       # 2 launches instead of 6 batched GEMMs
    assert rel(got_out, want_out) <= 2e-3
    gmax = max(float(np.abs(g).max()) for g in want_g)
    for g, w in zip(got_g, want_g):
        # (key.bias has a mathematically zero gradient: both paths return rounding noise of ~4096 summed rows there)
        assert float(np.abs(g - w).max()) <= 5e-3 * max(float(np.abs(w).max()), 2e-3 * gmax)
