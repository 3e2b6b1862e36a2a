"""MNIST MLP training step on lightgrad_b200 (model and loop of the reference's examples/mnist.py:24-67).

The dataset download of the reference (data.py, needs network) is out of scope: batches are
synthetic tensors of the MNIST shapes.
"""
import numpy as np
import lightgrad_b200 as light
import lightgrad_b200.nn as nn


class NN(nn.Module):
    """784-128-10, no biases (examples/mnist.py:24-32)."""

    def __init__(self):
        nn.Module.__init__(self)
        self.l1 = nn.Linear(28 * 28, 128, bias=False)
        self.l2 = nn.Linear(128, 10, bias=False)

    def forward(self, x):
        y = self.l1(x.reshape(-1, 28 * 28)).relu()
        return self.l2(y)


class CNN(nn.Module):
    """Two 3x3 convolutions with max-pooling and a linear head (examples/mnist.py:12-22)."""

    def __init__(self):
        nn.Module.__init__(self)
        self.c1 = nn.Conv2d(1, 8, kernelsize=3, bias=False, pad=0)
        self.c2 = nn.Conv2d(8, 16, kernelsize=3, bias=False, pad=0)
        self.l1 = nn.Linear(5 * 5 * 16, 10)

    def forward(self, x):
        y = self.c1(x).max_pool().relu()
        y = self.c2(y).max_pool().relu()
        return self.l1(y.reshape(-1, 5 * 5 * 16))


def synthetic_batch(batch=64, seed=0):
    rs = np.random.RandomState(seed)
    x = rs.uniform(0, 1, size=(batch, 1, 28, 28)).astype(np.float32)
    y = rs.randint(0, 10, size=(batch,)).astype(np.int16)
    return x, y


def train_step(model, optimizer, x, labels, T):
    """One step of the reference loop (mnist.py:54-64): one-hot by fancy assignment, mse, update."""
    y = model(x)
    one_hot = T.zeros((x.shape[0], 10))
    one_hot[range(x.shape[0]), labels] = 1
    loss = light.loss.mse(y, one_hot)
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return loss
