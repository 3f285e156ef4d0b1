"""CPU negative sampler with the reference's rejection rules (drop-in for unirec/data/transform/addnegsamples.py:15-115).
Used by the evaluation loaders and the CPU tests; training batches are built on the GPU (unirec_b200/data/device_loader.py)."""
import copy
import random

import numpy as np

from unirec_b200.data.history import UserHistoryCSR
from unirec_b200.utils.sampling import popularity_weights, prepare_aliased_randomizer


class AddNegSamples(object):
    MAX_TRIES = 100

    def __init__(self, n_users, n_items, n_neg, **kwargs):
        self.n_users, self.n_items, self.n_neg = n_users, n_items, n_neg
        hist = kwargs.get('user2history')
        if hist is not None and not isinstance(hist, UserHistoryCSR):
            hist = UserHistoryCSR.from_object_array(hist, n_users)
        self.history = hist
        self._sets = {}
        self.item_sampler = None
        pop = kwargs.get('item_popularity')
        if pop is not None:
            alpha = kwargs.get('neg_by_pop_alpha')
            self.item_sampler = prepare_aliased_randomizer(popularity_weights(pop, 1.0 if alpha is None else alpha))

    def _history_set(self, user_id):
        if self.history is None:
            return None
        s = self._sets.get(user_id)
        if s is None:
            s = self._sets[user_id] = set(self.history.history(int(user_id)).tolist())
        return s

    def _draw(self):
        return random.randint(1, self.n_items - 1) if self.item_sampler is None else self.item_sampler()

    def __call__(self, sample):
        """sample: object row [user_id, item_id(s), (label), ...] -> item ids become [positives..., negatives...]; a draw that
        fails MAX_TRIES times yields id 0; labels (when present) are extended with zeros."""
        sample = copy.deepcopy(sample)
        pos = sample[1]
        pos_list = list(pos) if isinstance(pos, (list, np.ndarray)) else [pos]
        banned = set(int(x) for x in pos_list)
        hist = self._history_set(sample[0])
        out = np.zeros(len(pos_list) + self.n_neg, dtype=int)
        out[:len(pos_list)] = pos_list
        for k in range(self.n_neg):
            for _ in range(self.MAX_TRIES):
                cand = int(self._draw())
                if cand not in banned and (hist is None or cand not in hist):
                    out[len(pos_list) + k] = cand
                    break
        sample[1] = out
        if len(sample) >= 3:
            labels = np.zeros(len(out), dtype=np.int32)
            labels[:len(pos_list)] = sample[2]
            sample[2] = labels
        return sample
