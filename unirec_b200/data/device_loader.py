"""DeviceBatchLoader: training batches built on the GPU.

The reference feeds the trainer from DataLoader worker processes that run a per-sample Python negative sampler
(0.4-1.3 ms per sample at K=256-1024, i.e. 3-10k samples/s in total).  Here the interaction columns, the CSR user history and the
alias table live in HBM; one `ur_build_batch` launch per step produces the whole batch contract on the device.  Iteration order,
batch size, shuffling and `len()` follow torch's DataLoader; `.dataset.return_key_2_index` is what the trainer reads."""
import numpy as np
import torch

from unirec_b200 import ops
from unirec_b200.data.history import UserHistoryCSR
from unirec_b200.utils.sampling import build_alias_table, popularity_weights


class DeviceBatchLoader:
    def __init__(self, dataset, batch_size, device, n_neg, n_users, n_items, max_seq_len=0, user_history=None,
                 history_mask_mode='unorder', seq_last=0, item_popularity=None, neg_by_pop_alpha=None, shuffle=False, seed=2022,
                 drop_last=False, rank=0, world=1):
        self.dataset = dataset
        self.batch_size, self.device, self.K, self.L = int(batch_size), torch.device(device), int(n_neg), int(max_seq_len)
        self.n_users, self.n_items = int(n_users), int(n_items)
        self.mask_mode, self.seq_last = history_mask_mode, int(seq_last)
        self.shuffle, self.seed, self.drop_last = bool(shuffle), int(seed), bool(drop_last)
        self.rank, self.world = int(rank), int(world)
        self.users = torch.from_numpy(np.ascontiguousarray(dataset.user_id)).to(self.device)
        self.items = torch.from_numpy(np.ascontiguousarray(dataset.item_id)).to(self.device)
        self.hist = None
        if user_history is not None:
            if not isinstance(user_history, UserHistoryCSR):
                user_history = UserHistoryCSR.from_object_array(user_history, self.n_users)
            self.hist = user_history.device_tensors(self.device)
        self.alias = None
        if item_popularity is not None:
            prob, alias = build_alias_table(popularity_weights(item_popularity, 1.0 if neg_by_pop_alpha is None else neg_by_pop_alpha))
            self.alias = (torch.from_numpy(prob).to(self.device), torch.from_numpy(alias).to(self.device))
        self.epoch, self.step = 0, 0
        self.with_seq = 'item_seq' in dataset.return_key_2_index

    def _my_count(self):
        """Samples per rank per epoch -- IDENTICAL on every rank: the stream is padded by wrapping around to a multiple of the world
        size (accelerate's even_batches behaviour, unirec/facility/trainer.py:261), because the row-sharded step issues fixed-shape
        collectives and every rank must run the same number of steps with the same batch sizes."""
        n = len(self.dataset)
        return (n + self.world - 1) // self.world

    def __len__(self):
        n = self._my_count()
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.dataset)
        if self.shuffle:
            g = torch.Generator(device='cpu').manual_seed(self.seed + self.epoch)
            order = torch.randperm(n, generator=g).to(self.device)
        else:
            order = torch.arange(n, device=self.device)
        pad = self._my_count() * self.world - n
        if pad:                                        # wrap around: every rank gets the same number of samples
            order = torch.cat([order, order[:pad]])
        order = order[self.rank::self.world]          # per-rank sharding of the sample stream (accelerate's prepared loader)
        self.epoch += 1
        ptr, hitems, hsorted = self.hist if self.hist is not None else (None, None, None)
        prob, alias = self.alias if self.alias is not None else (None, None)
        for i in range(len(self)):
            idx = order[i * self.batch_size:(i + 1) * self.batch_size]
            user_id, pos = self.users[idx], self.items[idx]
            item_id, label, item_seq, seq_len = ops.build_batch(
                user_id, pos, self.n_users, self.n_items, self.K, self.L if self.with_seq else 0, ptr, hitems, hsorted, prob, alias,
                self.mask_mode, self.seq_last, self.seed, self.step)
            self.step += 1
            out = [user_id, item_id, label]
            if self.with_seq:
                out += [item_seq, seq_len]
            yield tuple(out)
