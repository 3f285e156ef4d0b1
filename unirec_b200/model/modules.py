"""Parameter containers for the Transformer encoder.

The arithmetic of the reference's MultiHeadAttention / FeedForward / TransformerLayer / TransformerEncoder
(unirec/model/modules.py:247-433) runs in the CUDA engine (unirec_b200/engine.py: SASRecTower); these modules only
create and name the parameters, in the reference's order, so that (a) `state_dict` keys match
(`trm_encoder.layer.{i}.multi_head_attention.query.weight`, ...), and (b) a given seed yields bit-identical
initial weights (same constructor sequence -> same RNG consumption; SURVEY hard part H7).
"""
import copy

import torch.nn as nn

ACTIVATIONS = ('gelu', 'relu', 'swish', 'tanh', 'sigmoid')


class AttentionParams(nn.Module):
    def __init__(self, n_heads, hidden_size, hidden_dropout_prob, attn_dropout_prob, layer_norm_eps):
        super().__init__()
        if hidden_size % n_heads != 0:
            raise ValueError('The hidden size (%d) is not a multiple of the number of attention heads (%d)'
                             % (hidden_size, n_heads))
        self.num_attention_heads = n_heads
        self.attention_head_size = hidden_size // n_heads
        self.query = nn.Linear(hidden_size, hidden_size)
        self.key = nn.Linear(hidden_size, hidden_size)
        self.value = nn.Linear(hidden_size, hidden_size)
        self.attn_dropout = nn.Dropout(attn_dropout_prob)
        self.dense = nn.Linear(hidden_size, hidden_size)
        self.LayerNorm = nn.LayerNorm(hidden_size, eps=layer_norm_eps)
        self.out_dropout = nn.Dropout(hidden_dropout_prob)


class FeedForwardParams(nn.Module):
    def __init__(self, hidden_size, inner_size, hidden_dropout_prob, hidden_act, layer_norm_eps):
        super().__init__()
        if hidden_act not in ACTIVATIONS:
            raise KeyError(hidden_act)
        self.hidden_act = hidden_act
        self.dense_1 = nn.Linear(hidden_size, inner_size)
        self.dense_2 = nn.Linear(inner_size, hidden_size)
        self.LayerNorm = nn.LayerNorm(hidden_size, eps=layer_norm_eps)
        self.dropout = nn.Dropout(hidden_dropout_prob)


class TransformerLayerParams(nn.Module):
    def __init__(self, n_heads, hidden_size, inner_size, hidden_dropout_prob, attn_dropout_prob, hidden_act, layer_norm_eps):
        super().__init__()
        self.multi_head_attention = AttentionParams(n_heads, hidden_size, hidden_dropout_prob, attn_dropout_prob, layer_norm_eps)
        self.feed_forward = FeedForwardParams(hidden_size, inner_size, hidden_dropout_prob, hidden_act, layer_norm_eps)


class TransformerEncoderParams(nn.Module):
    """n_layers deep copies of ONE constructed layer, exactly as the reference builds it (modules.py:411-414)."""

    def __init__(self, n_layers=2, n_heads=2, hidden_size=64, inner_size=256, hidden_dropout_prob=0.5,
                 attn_dropout_prob=0.5, hidden_act='gelu', layer_norm_eps=1e-12):
        super().__init__()
        proto = TransformerLayerParams(n_heads, hidden_size, inner_size, hidden_dropout_prob, attn_dropout_prob,
                                       hidden_act, layer_norm_eps)
        self.layer = nn.ModuleList([copy.deepcopy(proto) for _ in range(n_layers)])
