"""SeqRecDataset (drop-in for unirec/data/dataset/seqrecdataset.py:20-68): BaseDataset + the user's history, cut by the history
transform and left-padded to max_seq_len (int32), plus its length."""
import numpy as np

from unirec_b200.constants.protocols import DataFileFormat
from .basedataset import BaseDataset


class SeqRecDataset(BaseDataset):
    def __init__(self, config, path, filename, transform=None):
        super().__init__(config, path, filename, transform)
        self.add_seq_transform = None

    def set_return_column_index(self):
        super().set_return_column_index()
        self.return_key_2_index['item_seq'] = len(self.return_key_2_index)
        self.return_key_2_index['item_seq_len'] = len(self.return_key_2_index)

    def add_user_history_transform(self, transform):
        self.add_seq_transform = transform

    def _padding(self, x):
        k = self.config['max_seq_len']
        res = np.zeros((k,), dtype=np.int32)
        n = min(len(x), k)
        if n:
            res[k - n:] = x[len(x) - n:]
        return res

    def __getitem__(self, index):
        elements = super().__getitem__(index)
        if self.config['data_format'] == DataFileFormat.T1_1.value:
            seq, seq_len, _ = self.add_seq_transform((elements[0], elements[1], elements[3]))
        else:
            seq, seq_len, _ = self.add_seq_transform((elements[0], elements[1]))
        return elements + (self._padding(seq), min(seq_len, self.config['max_seq_len']))
