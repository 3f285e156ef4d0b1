"""AvgHist (drop-in for unirec/model/sequential/avghist.py:9-55): user = (len+1)^-alpha * sum of history rows.
With `asymmetric` the history side reads a separate deep-copied table (`item_dst_embedding`), targets read
`item_src_embedding` (= `item_embedding`)."""
import copy

from .seqrec_base import SeqRecBase


class AvgHist(SeqRecBase):
    _tower_kind = 'avghist'

    def __init__(self, config):
        self.asymmetric = config['asymmetric']
        self.alpha = config['user_sequence_alpha']
        super().__init__(config)

    def _define_model_layers(self):
        self.item_src_embedding = self.item_embedding
        self.item_dst_embedding = copy.deepcopy(self.item_embedding) if self.asymmetric else self.item_embedding
