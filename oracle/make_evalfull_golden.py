"""Generate tests/golden_eval/one_vs_all_*.npz from the UNMODIFIED reference evaluator (TEST INFRASTRUCTURE ONLY).

    python oracle/make_evalfull_golden.py        # in the build container, where /root/reference exists

The reference model classes (SASRec with item + user bias and tau; MF) are built on the CPU under a fixed seed and pushed through
`OnePositiveEvaluator.evaluate_with_full_items` (unirec/facility/evaluation/evaluator_abc.py:190-278) with a user history that holds
duplicates, the target itself, a user without history and a user id beyond the history array.  Saved: parameters, batches, the
history (CSR), the per-sample metric vectors and the ranks (recovered exactly from group_auc = (V-1-rank)/(V-1)).
Shims (none touches arithmetic): `np.Inf` (numpy 2), a single-process stand-in for the accelerate.Accelerator surface the evaluator
uses (`unwrap_model`, `is_local_main_process`, `gather_for_metrics`, `device`).
"""
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get('UNIREC_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden_eval')
METRICS = "['hit@1;5;10', 'ndcg@5;10', 'mrr@5', 'mrr', 'ndcg', 'group_auc']"

CASES = {
    'one_vs_all_sasrec_bias': dict(model='SASRec', n_items=311, n_users=41, embedding_size=32, hidden_size=32, n_layers=1, n_heads=2,
                                   inner_size=64, max_seq_len=8, loss_type='softmax', has_item_bias=True, has_user_bias=True,
                                   tau=0.7, init_std=0.3, B=9, n_batches=2),
    'one_vs_all_mf': dict(model='MF', n_items=257, n_users=53, embedding_size=64, loss_type='bpr', init_std=0.3, B=16, n_batches=2),
}


class FakeAccelerator:
    device = torch.device('cpu')
    is_local_main_process = True

    def unwrap_model(self, m):
        return m

    def gather_for_metrics(self, t):
        return t


class FakeLoader:
    def __init__(self, batches, key2index, cfg):
        self.batches = batches

        class DS:
            pass
        self.dataset = DS()
        self.dataset.return_key_2_index = key2index
        self.dataset.config = cfg

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        return iter(self.batches)


def main():
    sys.path.insert(0, REF)
    if not hasattr(np, 'Inf'):
        np.Inf = np.inf
    from unirec.facility.evaluation.onepos import OnePositiveEvaluator
    from unirec.utils import argument_parser, general
    os.makedirs(OUT, exist_ok=True)
    saved_argv, sys.argv = sys.argv, sys.argv[:1]
    for name, case in CASES.items():
        case = dict(case)
        B, nb = case.pop('B'), case.pop('n_batches')
        args = dict(dataset='example', exp_name='golden', train_file_format='user-item', hidden_dropout_prob=0.0,
                    attn_dropout_prob=0.0, scheduler='none')
        args.update(case)
        cfg = argument_parser.parse_arguments(args)
        cfg['device'] = torch.device('cpu')
        general.init_seed(2022)
        model = general.get_class_instance(cfg['model'], 'unirec/model')(cfg)
        model.eval()
        V, U, L = cfg['n_items'], cfg['n_users'], int(cfg.get('max_seq_len', 0) or 0)
        gen = torch.Generator().manual_seed(99)
        rng = np.random.RandomState(5)
        # user history: object array like general.load_user_history returns (entries may be None)
        hist = np.empty(U - 3, dtype=object)            # users U-3.. have ids beyond the array
        for u in range(1, U - 3):
            n = int(rng.randint(0, 12))
            h = rng.randint(1, V, size=n).astype(np.int64)
            if n >= 4:
                h[1] = h[0]                              # duplicate entry
            hist[u] = h if n else None
        batches = []
        seq_model = cfg['model'] != 'MF'
        for b in range(nb):
            user_id = torch.randint(1, U, (B,), generator=gen, dtype=torch.int64)
            item_id = torch.randint(1, V, (B,), generator=gen, dtype=torch.int64)
            # target inside the user's own history (evaluator_abc.py:251-257 masks it and restores its score at slot 0)
            for k in range(0, B, 3):
                u = int(user_id[k])
                if u < len(hist) and hist[u] is not None:
                    item_id[k] = int(hist[u][-1])
            if seq_model:
                lens = torch.randint(1, L + 1, (B,), generator=gen)
                seq = torch.zeros(B, L, dtype=torch.int32)
                for i in range(B):
                    n = int(lens[i])
                    seq[i, L - n:] = torch.randint(1, V, (n,), generator=gen).to(torch.int32)
                batches.append((user_id, item_id, seq, lens.to(torch.int64)))
            else:
                batches.append((user_id, item_id))
        key2index = {'user_id': 0, 'item_id': 1, 'item_seq': 2, 'item_seq_len': 3} if seq_model else {'user_id': 0, 'item_id': 1}
        ds_cfg = dict(cfg)
        ds_cfg['data_format'] = 'user-item'
        ev = OnePositiveEvaluator(METRICS, -1, cfg, FakeAccelerator())
        per_batch = []
        merge = ev.merge_scores
        ev.merge_scores = lambda all_results: (per_batch.extend(all_results), merge(all_results))[1]   # keep the per-sample vectors
        merged = ev.evaluate_with_full_items(FakeLoader(batches, key2index, ds_cfg), model, hist)
        res = {k: np.concatenate([np.asarray(r[k]).reshape(-1) for r in per_batch]) for k in per_batch[0]}
        for k, v in merged.items():
            res['merged_' + k] = v
        out = {}
        for k, v in model.state_dict().items():
            out['param/' + k] = v.detach().numpy().copy()
        for k, v in res.items():
            out['metric/' + k] = np.asarray(v, dtype=np.float64)
        auc = np.asarray(res['group_auc'], dtype=np.float64)
        out['rank'] = np.rint((V - 1) * (1.0 - auc)).astype(np.int64)
        for i, b in enumerate(batches):
            for kname, idx in key2index.items():
                out['batch%d/%s' % (i, kname)] = b[idx].numpy()
        ptr = np.zeros(len(hist) + 1, dtype=np.int64)
        items = []
        for u in range(len(hist)):
            h = hist[u] if hist[u] is not None else np.zeros(0, np.int64)
            ptr[u + 1] = ptr[u] + len(h)
            items.append(h)
        out['hist_ptr'] = ptr
        out['hist_items'] = np.concatenate(items).astype(np.int32)
        keep = {k: v for k, v in cfg.items() if isinstance(v, (int, float, str, bool)) and k not in ('exp_name', 'config_dir')}
        keep['B'], keep['n_batches'], keep['metrics'] = B, nb, METRICS
        out['config_json'] = np.frombuffer(json.dumps(keep, sort_keys=True).encode(), dtype=np.uint8)
        path = os.path.join(OUT, name + '.npz')
        np.savez_compressed(path, **out)
        print('%-26s ranks %s  %.1f KB' % (name, out['rank'][:8], os.path.getsize(path) / 1024))
    sys.argv = saved_argv


if __name__ == '__main__':
    main()
