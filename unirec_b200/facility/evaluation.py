"""One-positive ranking evaluation on the device (replaces the CPU/numpy path of
unirec/facility/evaluation/evaluator_abc.py:124-278 and onepos.py:104-175 for the protocols the hot path uses).

one_vs_k  : the batch carries [B,1+K] candidates, positive first; rank = #negatives scoring above the positive.
one_vs_all: rank of the target among ALL items, computed in item tiles (no [B,V] matrix is kept), items in the
            user's history and the padding id excluded (reference masks them with NINF, evaluator_abc.py:249-265).
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

    def _history_tensor(self, user_ids, device):
        uh = self.user_history
        rows = [np.asarray(uh[int(u)]) if uh is not None and int(u) < len(uh) and uh[int(u)] is not None else np.zeros(0, np.int64)
                for u in user_ids.tolist()]
        width = max(1, max(len(r) for r in rows))
        out = np.zeros((len(rows), width), dtype=np.int64)
        for i, r in enumerate(rows):
            out[i, :len(r)] = r
        return torch.from_numpy(out).to(device)

    @torch.no_grad()
    def _ranks(self, model, samples):
        if self.protocol == EvaluationProtocal.OneVSK.value:
            _, scores, _, _ = model(**samples)
            if self.group_size > 0:
                scores = scores.reshape(-1, self.group_size)
            pos = scores[:, :1]
            return (scores[:, 1:] > pos).sum(1), scores.shape[1]
        # one_vs_all
        keys = inspect.signature(model.forward_user_emb).parameters
        user = model.forward_user_emb(**{k: v for k, v in samples.items() if k in keys})
        table = model.forward_all_item_emb(numpy=False)
        bias = model.item_bias.data if model.has_item_bias else None
        target = samples['item_id'].reshape(samples['item_id'].shape[0], -1)[:, 0].long()
        t_emb = table[target]
        pos = (user * t_emb).sum(-1, keepdim=True)
        if bias is not None:
            pos = pos + bias[target].unsqueeze(1)
        V = table.shape[0]
        greater = torch.zeros(user.shape[0], dtype=torch.int64, device=user.device)
        for s in range(0, V, self.item_tile):
            e = min(V, s + self.item_tile)
            sc = user @ table[s:e].t()
            if bias is not None:
                sc = sc + bias[s:e]
            greater += (sc > pos).sum(1)
        hist = self._history_tensor(samples['user_id'], user.device) if 'user_id' in samples else None
        excl = torch.zeros_like(greater)
        pad_sc = (user * table[0]).sum(-1, keepdim=True) + (bias[0] if bias is not None else 0.0)
        excl += (pad_sc > pos).squeeze(1).long()
        if hist is not None:
            hs = torch.einsum('bd,bhd->bh', user, table[hist])
            if bias is not None:
                hs = hs + bias[hist]
            live = (hist != 0) & (hist != target.unsqueeze(1))
            # a history item listed twice must be excluded once: keep the first occurrence only
            srt, _ = torch.sort(hist, dim=1)
            dup_total = ((srt[:, 1:] == srt[:, :-1]) & (srt[:, 1:] != 0)).sum(1)
            cnt = ((hs > pos) & live).sum(1)
            if int(dup_total.sum()) > 0:
                first = torch.ones_like(hist, dtype=torch.bool)
                for b in torch.nonzero(dup_total).flatten().tolist():
                    seen = set()
                    for j, it in enumerate(hist[b].tolist()):
                        first[b, j] = it not in seen
                        seen.add(it)
                cnt = ((hs > pos) & live & first).sum(1)
            excl += cnt
        return greater - excl, V - 1

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
