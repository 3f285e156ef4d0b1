"""User interaction histories in CSR form ("CSR-ify user2history once instead of ndarray[object]").

`ptr[u] .. ptr[u+1]` delimits user u's items in chronological order; `sorted_items` holds the same slices sorted ascending for
O(log n) membership tests in the device-side negative sampler.  Built from the reference's representation (an object ndarray of
per-user item arrays, unirec/utils/general.py:111-149) or straight from a DataFrame."""
import numpy as np
import torch


class UserHistoryCSR:
    def __init__(self, ptr, items):
        self.ptr = np.asarray(ptr, dtype=np.int64)
        self.items = np.asarray(items, dtype=np.int32)
        self.n_users = len(self.ptr) - 1
        self._sorted = None
        self._dev = {}

    @classmethod
    def from_object_array(cls, user2history, n_users=None):
        n = len(user2history) if n_users is None else max(n_users, len(user2history))
        lens = np.zeros(n, dtype=np.int64)
        for u, h in enumerate(user2history):
            if h is not None:
                lens[u] = len(h)
        ptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lens, out=ptr[1:])
        items = np.zeros(int(ptr[-1]), dtype=np.int32)
        for u, h in enumerate(user2history):
            if h is not None and len(h):
                items[ptr[u]:ptr[u + 1]] = np.asarray(h, dtype=np.int64)
        return cls(ptr, items)

    @classmethod
    def from_interactions(cls, user_ids, item_ids, n_users):
        """Group (user, item) rows by user, keeping file order inside each user (reference: groupby('user_id'))."""
        user_ids = np.asarray(user_ids, dtype=np.int64)
        order = np.argsort(user_ids, kind='stable')
        counts = np.bincount(user_ids, minlength=n_users)
        ptr = np.zeros(n_users + 1, dtype=np.int64)
        np.cumsum(counts, out=ptr[1:])
        return cls(ptr, np.asarray(item_ids)[order])

    def history(self, u):
        if u < 0 or u >= self.n_users:
            return self.items[:0]
        return self.items[self.ptr[u]:self.ptr[u + 1]]

    def to_object_array(self):
        res = np.empty(self.n_users, dtype=object)
        for u in range(self.n_users):
            if self.ptr[u + 1] > self.ptr[u]:
                res[u] = self.items[self.ptr[u]:self.ptr[u + 1]].astype(np.int64)
        return res

    @property
    def sorted_items(self):
        if self._sorted is None:
            out = self.items.copy()
            # sort every slice: one global sort on (user, item) keys
            owner = np.repeat(np.arange(self.n_users, dtype=np.int64), np.diff(self.ptr))
            order = np.lexsort((out, owner))
            self._sorted = out[order]
        return self._sorted

    def device_tensors(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = (torch.from_numpy(self.ptr).to(device), torch.from_numpy(self.items).to(device),
                              torch.from_numpy(self.sorted_items).to(device))
        return self._dev[key]
