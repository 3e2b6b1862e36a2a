"""BERT (masked-LM head) on lightgrad_b200 -- the model of the reference's examples/bert.py:12-228.

Same module tree and parameter names as the reference (so a reference parameter dict loads with
``load_parameters``), same arithmetic.  Differences, all on the step path the north star names:
  * ``Embedding`` stays on the device and is differentiable (the reference round-trips through the
    CPU tensor and detaches the table, bert.py:19-21);
  * gelu / softmax(QK^T/sqrt(d)) / LayerNorm / Linear use the backend's fused operators when the
    tensor class provides them, otherwise the reference's compositions;
  * a training step exists (``train_step``): the reference file is inference-only (bert.py:332-356).
Tokenizer and ``from_pretrained`` (network access) are out of scope.
"""
import math
import numpy as np
import lightgrad_b200 as light
import lightgrad_b200.nn as nn
from lightgrad_b200 import loss as losses

BERT_BASE = dict(hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                 vocab_size=30522, max_position_embeddings=512, type_vocab_size=2,
                 attention_probs_dropout_prob=0.0, hidden_dropout_prob=0.0)


def gelu(x):
    """tanh-GELU; one fused kernel on the cuda backend, the reference's 9-op lambda elsewhere (bert.py:12)."""
    if hasattr(x, 'gelu'):
        return x.gelu()
    return 0.5 * x * (1.0 + (x * 0.7978845608 * (1.0 + 0.044715 * x * x)).tanh())


class BertEmbedding(nn.Module):
    def __init__(self, hidden_size, vocab_size, max_position_embeddings, type_vocab_size):
        nn.Module.__init__(self)
        self.word_embeddings = nn.Embedding(hidden_size, vocab_size)
        self.position_embeddings = nn.Embedding(hidden_size, max_position_embeddings)
        self.token_type_embeddings = nn.Embedding(hidden_size, type_vocab_size)
        self.LayerNorm = nn.LayerNorm(hidden_size)
        self._pos_cache = {}

    def forward(self, input_ids, token_type_ids=None):
        T = input_ids.__class__
        key = (T, input_ids.shape)
        if key not in self._pos_cache:
            # index tensors depend only on the batch shape: build them once, not every step
            self._pos_cache[key] = (
                T.zeros(input_ids.shape, dtype=np.int32, requires_grad=False),
                T.from_numpy(np.arange(input_ids.shape[-1], dtype=np.int32), requires_grad=False))
        zeros, position_ids = self._pos_cache[key]
        if token_type_ids is None:
            token_type_ids = zeros
        embedd = self.word_embeddings(input_ids) + self.position_embeddings(position_ids) \
            + self.token_type_embeddings(token_type_ids)
        return self.LayerNorm(embedd)


class BertSelfAttention(nn.Module):
    def __init__(self, hidden_size, num_attention_heads):
        nn.Module.__init__(self)
        assert hidden_size % num_attention_heads == 0
        self.h = num_attention_heads
        self.d = hidden_size // num_attention_heads
        self.query = nn.Linear(hidden_size, hidden_size)
        self.key = nn.Linear(hidden_size, hidden_size)
        self.value = nn.Linear(hidden_size, hidden_size)

    def forward(self, hidden, attention_mask=None, need_probs=True):
        if attention_mask is None and not need_probs and hasattr(hidden, 'self_attention') \
                and len(hidden.shape) == 3:
            # whole block as one graph node: grouped QKV projection, per-head views, stacked gradients
            q, k, v = self.query, self.key, self.value
            return hidden.self_attention(q.weight, q.bias, k.weight, k.bias, v.weight, v.bias, heads=self.h), None
        Q, K, V = self.query(hidden), self.key(hidden), self.value(hidden)
        b, s, _ = K.shape
        Q = Q.reshape(b, s, self.h, self.d).transpose(0, 2, 1, 3)
        K = K.reshape(b, s, self.h, self.d).transpose(0, 2, 3, 1)
        V = V.reshape(b, s, self.h, self.d).transpose(0, 2, 1, 3)
        if attention_mask is None and getattr(Q.__class__, 'has_scaled_softmax', False):
            scores = (Q @ K).softmax(axis=-1, scale=1.0 / math.sqrt(self.d))
        else:
            scores = Q @ K / math.sqrt(self.d)
            if attention_mask is not None:
                attention_mask = attention_mask.reshape(1, 1, *attention_mask.shape)
                attention_mask = (1.0 - attention_mask) * -10000.0
                scores = scores + attention_mask.detach()
            scores = scores.softmax(axis=-1)
        if getattr(scores.__class__, 'dot_supports_layout', False):
            # the product lands in V's (batch, seq, head, dim) memory order: the head merge below is a view
            context = scores.dot(V, out_layout='rhs').transpose(0, 2, 1, 3)
        else:
            context = (scores @ V).transpose(0, 2, 1, 3)
        return context.reshape(b, s, self.h * self.d), scores


class BertAttention(nn.Module):
    def __init__(self, hidden_size, num_attention_heads):
        nn.Module.__init__(self)
        self.self = BertSelfAttention(hidden_size, num_attention_heads)
        self.output = nn.Module()
        self.output.dense = nn.Linear(hidden_size, hidden_size)
        self.output.LayerNorm = nn.LayerNorm(hidden_size)

    def forward(self, hidden_in, attention_mask=None, need_probs=True):
        hidden, attentions = self.self(hidden_in, attention_mask=attention_mask, need_probs=need_probs)
        hidden = self.output.dense(hidden)
        return self.output.LayerNorm(hidden, residual=hidden_in), attentions


class BertLayer(nn.Module):
    def __init__(self, hidden_size, intermediate_size, num_attention_heads):
        nn.Module.__init__(self)
        self.attention = BertAttention(hidden_size, num_attention_heads)
        self.intermediate = nn.Module()
        self.intermediate.dense = nn.Linear(hidden_size, intermediate_size)
        self.output = nn.Module()
        self.output.dense = nn.Linear(intermediate_size, hidden_size)
        self.output.LayerNorm = nn.LayerNorm(hidden_size)

    def mlp(self, x):
        if hasattr(x, 'mlp_gelu'):
            # both GEMMs, the bias adds and the GELU (forward and derivative) as one graph node
            d1, d2 = self.intermediate.dense, self.output.dense
            return x.mlp_gelu(d1.weight, d1.bias, d2.weight, d2.bias)
        return self.output.dense(gelu(self.intermediate.dense(x)))

    def forward(self, hidden, attention_mask=None, need_probs=True):
        hidden, attentions = self.attention(hidden, attention_mask, need_probs=need_probs)
        return self.output.LayerNorm(self.mlp(hidden), residual=hidden), attentions


class BertModel(nn.Module):
    def __init__(self, hidden_size, intermediate_size, num_hidden_layers, num_attention_heads, vocab_size,
                 max_position_embeddings, type_vocab_size, attention_probs_dropout_prob=0.0,
                 hidden_dropout_prob=0.0):
        nn.Module.__init__(self)
        self.embeddings = BertEmbedding(hidden_size, vocab_size, max_position_embeddings, type_vocab_size)
        self.encoder = nn.Module()
        self.encoder.layer = nn.ModuleList(*[
            BertLayer(hidden_size, intermediate_size, num_attention_heads) for _ in range(num_hidden_layers)])

    def forward(self, input_ids, attention_mask=None, token_type_ids=None):
        hidden = self.embeddings(input_ids, token_type_ids=token_type_ids)
        for layer in self.encoder.layer:
            # the attention probabilities are not returned from here: layers may fuse the whole block
            hidden, _ = layer(hidden, attention_mask=attention_mask, need_probs=False)
        return hidden


class BertForMaskedLM(nn.Module):
    def __init__(self, hidden_size, intermediate_size, num_hidden_layers, num_attention_heads, vocab_size,
                 max_position_embeddings, type_vocab_size, attention_probs_dropout_prob=0.0,
                 hidden_dropout_prob=0.0, **kwargs):
        nn.Module.__init__(self)
        self.bert = BertModel(hidden_size=hidden_size, intermediate_size=intermediate_size,
                              num_hidden_layers=num_hidden_layers, num_attention_heads=num_attention_heads,
                              vocab_size=vocab_size, max_position_embeddings=max_position_embeddings,
                              type_vocab_size=type_vocab_size)
        self.cls = nn.Module()
        self.cls.predictions = nn.Module()
        self.cls.predictions.transform = nn.Module()
        self.cls.predictions.transform.dense = nn.Linear(hidden_size, hidden_size)
        self.cls.predictions.transform.LayerNorm = nn.LayerNorm(hidden_size)
        self.cls.predictions.decoder = nn.Linear(hidden_size, vocab_size, bias=False)
        self.cls.predictions.bias = nn.default_tensor().zeros(vocab_size)
        self.vocab_size = vocab_size

    def forward(self, input_ids, attention_mask=None, token_type_ids=None):
        h = self.bert(input_ids=input_ids, attention_mask=attention_mask, token_type_ids=token_type_ids)
        h = self.cls.predictions.transform.dense(h)
        h = gelu(h)
        h = self.cls.predictions.transform.LayerNorm(h)
        dec = self.cls.predictions.decoder
        if hasattr(h, 'linear'):
            # decoder bias lives outside the Linear (bert.py:209-210); fuse it into the GEMM epilogue
            return h.linear(dec.weight, self.cls.predictions.bias)
        return dec(h) + self.cls.predictions.bias


def train_step(model, optimizer, input_ids, labels):
    """One masked-LM training step: forward, cross entropy over all positions, backward, update.

    Returns the loss tensor (shape ()); nothing here synchronises with the device."""
    logits = model(input_ids)
    loss = losses.cross_entropy(logits.reshape(-1, model.vocab_size), labels)
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return loss


def synthetic_batch(batch, seq_len, vocab_size, seed=1):
    """Synthetic token ids / labels with the seeds SURVEY.md 8(d) config 4 names."""
    ids = np.random.RandomState(seed).randint(0, vocab_size, size=(batch, seq_len)).astype(np.int32)
    labels = np.random.RandomState(seed + 1).randint(0, vocab_size, size=(batch * seq_len,)).astype(np.int32)
    return ids, labels
