"""One-positive ranking evaluation on the device (replaces the CPU/numpy path of
unirec/facility/evaluation/evaluator_abc.py:124-278 and onepos.py:104-175 for the protocols the hot path uses).

one_vs_k  : the batch carries [B,1+K] candidates, positive first; rank = #negatives scoring above the positive.
one_vs_all: rank of the target among ALL items, counted on the device by csrc/evalrank.cu (no [B,V] matrix; history, padding id
            and the target's own slot masked with NINF like evaluator_abc.py:249-257; row-sharded tables: every rank counts over
            the rows it owns and the counts are summed).
Metrics follow the reference definitions for a single positive: hit@k = [rank<k], ndcg@k = [rank<k]/log2(rank+2),
mrr@k = [rank<k]/(rank+1), group_auc = fraction of candidates ranked below the positive.
"""
import inspect
from ast import literal_eval

import numpy as np
import torch
import torch.distributed as dist

from unirec_b200.constants.protocols import EvaluationProtocal


class RankEvaluator(object):
    def __init__(self, metrics_str, group_size=-1, config=None, accelerator=None, protocol='one_vs_k', user_history=None):
        self.metrics_list = literal_eval(metrics_str) if isinstance(metrics_str, str) else list(metrics_str)
        self.group_size = group_size
        self.config = config
        self.accelerator = accelerator
        self.protocol = protocol
        self.user_history = user_history
        self.item_tile = 1 << 16

    def _history_csr(self, device):
        """The users' histories as CSR device tensors (ptr, items, sorted), built once (unirec_b200/data/history.py)."""
        if self.user_history is None:
            return None
        key = str(device)
        cache = self.__dict__.setdefault('_hist_dev', {})
        if key not in cache:
            from unirec_b200.data.history import UserHistoryCSR
            uh = self.user_history
            if not isinstance(uh, UserHistoryCSR):
                uh = UserHistoryCSR.from_object_array(uh)
            cache[key] = uh.device_tensors(device)
        return cache[key]

    @torch.no_grad()
    def _ranks(self, model, samples):
        if self.protocol == EvaluationProtocal.OneVSK.value:
            _, scores, _, _ = model(**samples)
            if self.group_size > 0:
                scores = scores.reshape(-1, self.group_size)
            pos = scores[:, :1]
            return (scores[:, 1:] > pos).sum(1), scores.shape[1]
        # one_vs_all: rank of the target among all items, counted by ur_rank_count / ur_rank_exclude (csrc/evalrank.cu) over the rows
        # this rank owns -- no [B, V] score matrix, no host round trip, history / padding / target excluded on the device
        keys = inspect.signature(model.forward_user_emb).parameters
        user = model.forward_user_emb(**{k: v for k, v in samples.items() if k in keys})
        target = samples['item_id'].reshape(samples['item_id'].shape[0], -1)[:, 0].long().contiguous()
        user_id = samples['user_id'].long().contiguous() if 'user_id' in samples else None
        hist = self._history_csr(user.device) if user_id is not None else None
        counts = model._engine.rank_one_vs_all(user, target, user_id=user_id, hist=hist)
        return counts.long(), int(model.n_items)

    @torch.no_grad()
    def evaluate(self, data, model, verbose=0, predict_only=False):
        model.eval()
        key2index = data.dataset.return_key_2_index
        ranks, n_cand = [], 1
        for inter_data in data:
            samples = {k: inter_data[v] for k, v in key2index.items()}
            r, n_cand = self._ranks(model, samples)
            ranks.append(r)
        rank = torch.cat(ranks).double()
        if self.accelerator is not None and self.accelerator.distributed:
            rank = self.accelerator.gather_for_metrics(rank)
        return self.metrics_from_ranks(rank, n_cand)

    def metrics_from_ranks(self, rank, n_cand):
        res = {}
        for m in self.metrics_list:
            name, _, ks = m.partition('@')
            cutoffs = [int(k) for k in ks.split(';')] if ks else [None]
            for k in cutoffs:
                inside = torch.ones_like(rank) if k is None else (rank < k).double()
                if name == 'hit':
                    v = inside
                elif name == 'ndcg':
                    v = inside / torch.log2(rank + 2.0)
                elif name == 'mrr':
                    v = inside / (rank + 1.0)
                elif name == 'group_auc':
                    v = 1.0 - rank / max(n_cand - 1, 1)
                else:
                    raise ValueError('metric %r is not implemented by the device evaluator' % name)
                res[name if k is None else '%s@%d' % (name, k)] = float(v.mean())
        return res
