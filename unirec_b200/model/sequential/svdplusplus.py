"""SVD++ (drop-in for unirec/model/sequential/svdplusplus.py:11-39): user = U[user_id] + (len+1)^-alpha * sum of
history rows from a second item table."""
import copy

from .seqrec_base import SeqRecBase


class SVDPlusPlus(SeqRecBase):
    _tower_kind = 'svdpp'

    def __init__(self, config):
        self.alpha = config['user_sequence_alpha']
        super().__init__(config)

    def _define_model_layers(self):
        self.item_src_embedding = self.item_embedding
        self.item_dst_embedding = copy.deepcopy(self.item_embedding)
