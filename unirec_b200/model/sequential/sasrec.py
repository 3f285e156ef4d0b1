"""SASRec (drop-in for unirec/model/sequential/sasrec.py:10-76).

Parameters are created exactly as the reference does (position table with max_seq_len+1 rows, encoder layers as
deep copies, input LayerNorm); the forward/backward arithmetic is the CUDA engine's SASRecTower: fused
gather+position+LayerNorm, fused QKV projection, shared-memory attention with the additive -10000 mask (causal only
when use_position_emb), post-LN residual blocks, last position as the user vector.
"""
import torch.nn as nn

from unirec_b200.model import modules
from .seqrec_base import SeqRecBase


class SASRec(SeqRecBase):
    _tower_kind = 'sasrec'

    def __init__(self, config):
        self.n_layers = config['n_layers']
        self.n_heads = config['n_heads']
        self.inner_size = config['inner_size']
        self.hidden_dropout_prob = config['hidden_dropout_prob']
        self.attn_dropout_prob = config['attn_dropout_prob']
        self.hidden_act = config['hidden_act']
        self.layer_norm_eps = float(config['layer_norm_eps'])
        self.max_seq_len = config['max_seq_len']
        self.use_pos_emb = config['use_position_emb']
        super().__init__(config)

    def _define_model_layers(self):
        if self.hidden_size != self.embedding_size:
            raise ValueError('SASRec adds item and position embeddings: hidden_size must equal embedding_size')
        self.position_embedding = nn.Embedding(self.max_seq_len + 1, self.hidden_size) if self.use_pos_emb else None
        self.trm_encoder = modules.TransformerEncoderParams(
            n_layers=self.n_layers, n_heads=self.n_heads, hidden_size=self.hidden_size, inner_size=self.inner_size,
            hidden_dropout_prob=self.hidden_dropout_prob, attn_dropout_prob=self.attn_dropout_prob,
            hidden_act=self.hidden_act, layer_norm_eps=self.layer_norm_eps)
        self.LayerNorm = nn.LayerNorm(self.hidden_size, eps=self.layer_norm_eps)
        self.dropout = nn.Dropout(self.hidden_dropout_prob)
