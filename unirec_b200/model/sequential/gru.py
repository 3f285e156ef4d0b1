"""GRU4Rec-style tower (drop-in for unirec/model/sequential/gru.py:9-35).  nn.GRU / nn.Linear are kept as parameter
containers (torch's default GRU init is part of the seed contract); the recurrence runs in the engine's GRUTower."""
import torch.nn as nn

from .seqrec_base import SeqRecBase


class GRU(SeqRecBase):
    _tower_kind = 'gru'

    def _define_model_layers(self):
        self.num_layers = 1
        self.emb_dropout = nn.Dropout(self.dropout_prob)
        self.gru_layers = nn.GRU(input_size=self.embedding_size, hidden_size=self.hidden_size, num_layers=self.num_layers,
                                 bias=True, batch_first=True)
        self.dense = nn.Linear(self.hidden_size, self.embedding_size)
