"""History lookup + masking (drop-in for unirec/data/transform/adduserhistory.py:10-73, time sequences excluded)."""
import random

import numpy as np

from unirec_b200.constants.protocols import DataFileFormat, HistoryMaskMode
from unirec_b200.data.history import UserHistoryCSR


class AddUserHistory(object):
    def __init__(self, user2history, mask_mode='unorder', user2history_time=None, seq_last=0, data_format=None):
        if user2history is not None and not isinstance(user2history, UserHistoryCSR):
            user2history = UserHistoryCSR.from_object_array(user2history)
        self.history = user2history
        self.mask_mode, self.seq_last, self.data_format = mask_mode, seq_last, data_format
        self.empty_history = np.zeros((1,), dtype=np.int32)

    def __call__(self, sample):
        """sample = (user_id, item_id(s)[, max_len]) -> (history, len(history), None)."""
        items = sample[1]
        targets = set(int(x) for x in items) if isinstance(items, (list, np.ndarray)) else {int(items)}
        h = self.history.history(int(sample[0])) if self.history is not None else self.empty_history[:0]
        if len(h) == 0:
            h = self.empty_history
        if self.mask_mode == HistoryMaskMode.Unorder.value:
            h = np.where(np.isin(h, list(targets)), 0, h)
        elif self.mask_mode == HistoryMaskMode.Autoregressive.value:
            if self.data_format == DataFileFormat.T1_1.value:
                h = h[:sample[2]]
            else:
                hits = np.flatnonzero(np.isin(h, list(targets)))
                if len(hits):
                    h = h[:int(hits[-1] if self.seq_last else random.choice(hits.tolist()))]
        return h, len(h), None
