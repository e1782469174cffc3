"""Parameter containers for the XML encoders whose `forward` runs on the xmlb200 CUDA kernels.

They mirror the reference module tree (reference baselines/crossmodal_moment_localization/
model_components.py:67-89,141-163,201-216,244-317) only as far as checkpoint compatibility requires: identical
attribute names, parameter shapes and `state_dict` keys.  The arithmetic itself is in
tvretrieval_b200/csrc/*.cu.  In train mode (`module.training`, dropout p > 0) the reference's nn.Dropout layers are
applied by the counter-based xmlb_dropout kernel (ops.dropout) at the same places; in eval mode they are the
identity, exactly like the reference.
"""
import torch.nn as nn

from . import ops


def _drop(module, x):
    """module.dropout (an nn.Dropout, kept for its `p` and for state_dict/attribute parity) applied by ops.dropout."""
    p = module.p
    if not module.training or p <= 0:
        return x
    return ops.dropout(x, p, ops.new_seed())


class TrainablePositionalEncoding(nn.Module):
    """LN(x + E[0:L]) -- reference model_components.py:76-89."""

    def __init__(self, max_position_embeddings, hidden_size, dropout=0.1):
        super().__init__()
        self.position_embeddings = nn.Embedding(max_position_embeddings, hidden_size)
        self.LayerNorm = nn.LayerNorm(hidden_size)
        self.dropout = nn.Dropout(dropout)

    def forward(self, input_feat):
        seq_len = input_feat.shape[1]
        if seq_len > self.position_embeddings.num_embeddings:
            raise IndexError("sequence length %d exceeds the %d learned positions"
                             % (seq_len, self.position_embeddings.num_embeddings))
        out = ops.add_layernorm(input_feat, self.LayerNorm.weight, self.LayerNorm.bias,
                                add=self.position_embeddings.weight, add_rows=seq_len, eps=self.LayerNorm.eps)
        return _drop(self.dropout, out)


class LinearLayer(nn.Module):
    """relu(Linear(LN(x))) -- reference model_components.py:156-163; the Linear lives at `net.1`."""

    def __init__(self, in_hsz, out_hsz, layer_norm=True, dropout=0.1, relu=True):
        super().__init__()
        self.relu = relu
        self.layer_norm = layer_norm
        if layer_norm:
            self.LayerNorm = nn.LayerNorm(in_hsz)
        self.net = nn.Sequential(nn.Dropout(dropout), nn.Linear(in_hsz, out_hsz))

    def forward(self, x, precision=ops.DEFAULT_PRECISION):
        if self.layer_norm:
            x = ops.add_layernorm(x, self.LayerNorm.weight, self.LayerNorm.bias, eps=self.LayerNorm.eps)
        fc = self.net[1]
        return ops.linear(_drop(self.net[0], x), fc.weight, fc.bias, relu=self.relu, precision=precision)


class BertSelfAttention(nn.Module):
    """Q/K/V projections + masked multi-head attention -- reference model_components.py:266-303."""

    def __init__(self, config):
        super().__init__()
        if config.hidden_size % config.num_attention_heads != 0:
            raise ValueError("The hidden size (%d) is not a multiple of the number of attention heads (%d)"
                             % (config.hidden_size, config.num_attention_heads))
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = config.hidden_size // config.num_attention_heads
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        self.query = nn.Linear(config.hidden_size, self.all_head_size)
        self.key = nn.Linear(config.hidden_size, self.all_head_size)
        self.value = nn.Linear(config.hidden_size, self.all_head_size)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)

    def forward(self, query_states, key_states, value_states, attention_mask, precision=ops.DEFAULT_PRECISION):
        """attention_mask: (N, Lq or 1, L) float, 1 = attend."""
        q = ops.linear(query_states, self.query.weight, self.query.bias, precision=precision)
        k = ops.linear(key_states, self.key.weight, self.key.bias, precision=precision)
        v = ops.linear(value_states, self.value.weight, self.value.bias, precision=precision)
        p = self.dropout.p if self.training else 0.0
        return ops.attention(q, k, v, attention_mask, self.num_attention_heads, dropout_p=p,
                             seed=ops.new_seed() if p > 0 else 0)


class BertSelfOutput(nn.Module):
    """LN(dense(h) + x) -- reference model_components.py:313-317."""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor, precision=ops.DEFAULT_PRECISION):
        if self.training and self.dropout.p > 0:  # dense -> dropout -> + residual -> LN
            h = _drop(self.dropout, ops.linear(hidden_states, self.dense.weight, self.dense.bias,
                                               precision=precision))
            return ops.add_layernorm(h, self.LayerNorm.weight, self.LayerNorm.bias, add=input_tensor,
                                     eps=self.LayerNorm.eps)
        h = ops.linear(hidden_states, self.dense.weight, self.dense.bias, residual=input_tensor, precision=precision)
        return ops.add_layernorm(h, self.LayerNorm.weight, self.LayerNorm.bias, eps=self.LayerNorm.eps)


class BertAttention(nn.Module):
    """Self-attention block without FFN -- reference model_components.py:207-216."""

    def __init__(self, config):
        super().__init__()
        self.self = BertSelfAttention(config)
        self.output = BertSelfOutput(config)

    def forward(self, input_tensor, attention_mask, precision=ops.DEFAULT_PRECISION):
        att = self.self(input_tensor, input_tensor, input_tensor, attention_mask, precision=precision)
        return self.output(att, input_tensor, precision=precision)
