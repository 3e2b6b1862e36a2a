"""Loss functions with the reference's semantics (lightgrad/loss.py:4-24).

Kept on purpose: ``mse.backward`` is ``err * out_grad`` -- no 1/N factor -- and
``cross_entropy.backward`` is ``(softmax - onehot) / N * out_grad`` with no gradient for the labels.
``cross_entropy`` uses the backend's fused log-softmax + NLL kernels when the tensor class offers
them (one read of the logits forward, one read + one write backward) and otherwise the composed
softmax -> gather -> log -> mean path of the reference.
"""
from .autograd import Function


class mse(Function):
    """Mean Squared Error: mean((y - y_hat)^2) / 2."""

    def forward(ctx, y, y_hat):
        err = y - y_hat
        ctx.save_for_backward(err)
        return (err ** 2).mean() / 2

    def backward(ctx, out_grad):
        err, = ctx.get_saved_tensors()
        return err * out_grad


class cross_entropy(Function):
    """Softmax cross entropy over ``axis`` (labels are class indices)."""

    def forward(ctx, y, y_hat, axis=-1):
        fused = getattr(y.__class__, 'fused_cross_entropy', None)
        if fused is not None and len(y.shape) == 2 and axis in (-1, 1):
            loss, saved = fused(y, y_hat)
            ctx.save_for_backward(True, saved)
            return loss
        p = y.softmax(axis=axis)
        ctx.save_for_backward(False, (p, y_hat))
        return -p[range(y_hat.shape[0]), y_hat].log().mean()

    def backward(ctx, out_grad):
        fused, saved = ctx.get_saved_tensors()
        if fused:
            return saved[0].__class__.fused_cross_entropy_backward(saved, out_grad)
        p, y_hat = saved
        rows = range(y_hat.shape[0])
        # the reference mutates its saved softmax in place (loss.py:21-23); so do we
        p[rows, y_hat] = p[rows, y_hat] - 1
        p /= y_hat.shape[0]
        return p * out_grad
