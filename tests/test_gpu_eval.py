"""GPU: one-vs-all ranking kernels (csrc/evalrank.cu) against the reference evaluator's own output
(tests/golden_eval/one_vs_all_*.npz, produced by OnePositiveEvaluator.evaluate_with_full_items in oracle/make_evalfull_golden.py):
history masking incl. duplicates and target-in-history, users without history, item/user bias and tau, rank and metrics."""
import json
import os

import numpy as np
import pytest
import torch

from unirec_b200.utils import argument_parser, general

pytestmark = pytest.mark.gpu
DEV = 'cuda'
GDIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden_eval')


class _Loader:
    def __init__(self, batches, key2index):
        self.batches = batches

        class DS:
            pass
        self.dataset = DS()
        self.dataset.return_key_2_index = key2index

    def __iter__(self):
        return iter(self.batches)


def _load(name):
    z = np.load(os.path.join(GDIR, name + '.npz'))
    cfg = json.loads(bytes(z['config_json']).decode())
    args = dict(cfg)
    args.update(exp_name='ev', dataset='example')
    conf = argument_parser.parse_arguments(args, argv=[])
    conf['device'] = torch.device(DEV)
    general.init_seed(1)
    model = general.get_class_instance(conf['model'], 'unirec_b200/model')(conf).to(DEV)
    model.load_state_dict({k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('param/')})
    names = ['user_id', 'item_id', 'item_seq', 'item_seq_len'] if conf['model'] != 'MF' else ['user_id', 'item_id']
    batches = [tuple(torch.from_numpy(z['batch%d/%s' % (i, n)]).to(DEV) for n in names) for i in range(cfg['n_batches'])]
    ptr, items = z['hist_ptr'], z['hist_items']
    hist = np.empty(len(ptr) - 1, dtype=object)
    for u in range(len(hist)):
        if ptr[u + 1] > ptr[u]:
            hist[u] = items[ptr[u]:ptr[u + 1]].astype(np.int64)
    return z, cfg, conf, model, _Loader(batches, {n: i for i, n in enumerate(names)}), hist


@pytest.mark.parametrize('name', ['one_vs_all_sasrec_bias', 'one_vs_all_mf'])
def test_one_vs_all_ranks_and_metrics_match_reference_evaluator(name):
    from unirec_b200.facility.evaluation import RankEvaluator
    z, cfg, conf, model, loader, hist = _load(name)
    ev = RankEvaluator(cfg['metrics'], -1, conf, None, protocol='one_vs_all', user_history=hist)
    model.eval()
    ranks = []
    for b in loader:
        samples = {k: b[i] for k, i in loader.dataset.return_key_2_index.items()}
        r, n_cand = ev._ranks(model, samples)
        ranks.append(r.cpu().numpy())
        assert n_cand == cfg['n_items']
    assert np.array_equal(np.concatenate(ranks), z['rank'])               # integer work: exact
    res = ev.evaluate(loader, model)
    for k, v in res.items():
        assert abs(v - float(z['metric/merged_' + k])) < 1e-6, k


def test_sharded_counts_add_up_to_the_unsharded_rank():
    """Rank r of W counts over rows {id % W == r}; the partial target scores and counts sum to the single-table result."""
    from unirec_b200 import ops
    torch.manual_seed(3)
    V, d, S, W = 5003, 64, 77, 3
    table = torch.randn(V, d, device=DEV) * 0.3
    table[0] = 0
    user = torch.randn(S, d, device=DEV)
    target = torch.randint(1, V, (S,), device=DEV)
    ib = torch.randn(V, device=DEV) * 0.1
    uid = torch.randint(0, 40, (S,), device=DEV)
    lens = torch.randint(0, 9, (41,))
    ptr = torch.zeros(42, dtype=torch.int64)
    ptr[1:] = torch.cumsum(lens, 0)
    items = torch.randint(0, V, (int(ptr[-1]),), dtype=torch.int32)
    srt = items.clone()
    for u in range(41):
        srt[ptr[u]:ptr[u + 1]] = torch.sort(items[ptr[u]:ptr[u + 1]]).values
    ptr, srt = ptr.to(DEV), srt.to(DEV)

    def run(world):
        t = torch.zeros(S, device=DEV)
        c = torch.zeros(S, dtype=torch.int32, device=DEV)
        shards = [table[r::world].contiguous() for r in range(world)]
        for r in range(world):
            tr = torch.empty(S, device=DEV)
            ops.rank_target(shards[r], user, target, tr, item_bias=ib, user_id=uid, tau=0.5, world=world, rank=r)
            t += tr
        for r in range(world):
            ops.rank_count(shards[r], user, target, t, c, item_bias=ib, user_id=uid, tau=0.5, world=world, rank=r)
            ops.rank_exclude(shards[r], user, target, t, c, item_bias=ib, user_id=uid, tau=0.5, world=world, rank=r, hist_ptr=ptr,
                             hist_sorted=srt)
        return t, c

    t1, c1 = run(1)
    tw, cw = run(W)
    assert torch.equal(t1, tw) and torch.equal(c1, cw)
    # independent check in float64 (near-ties may flip a count by one)
    sc = (user.double() @ table.double().t() + ib.double()) / 0.5
    ts = sc.gather(1, target[:, None])
    mask = torch.ones_like(sc, dtype=torch.bool)
    mask[:, 0] = False
    mask[torch.arange(S), target] = False
    for s in range(S):
        u = int(uid[s])
        h = srt[ptr[u]:ptr[u + 1]].long()
        mask[s, h[h > 0]] = False
    ref = ((sc > ts) & mask).sum(1)
    assert int((ref - c1.long()).abs().max()) <= 1
    assert float((ref == c1.long()).float().mean()) > 0.95


def test_rank_count_at_catalogue_scale_matches_float64():
    from unirec_b200 import ops
    torch.manual_seed(5)
    V, d, S = 200_003, 128, 300
    table = torch.randn(V, d, device=DEV) * 0.1
    user = torch.randn(S, d, device=DEV)
    target = torch.randint(1, V, (S,), device=DEV)
    t = torch.empty(S, device=DEV)
    c = torch.zeros(S, dtype=torch.int32, device=DEV)
    ops.rank_target(table, user, target, t)
    ops.rank_count(table, user, target, t, c)
    ops.rank_exclude(table, user, target, t, c)
    ref = torch.zeros(S, dtype=torch.int64, device=DEV)
    ts = (user.double() * table[target].double()).sum(1, keepdim=True)
    for s0 in range(0, V, 50_000):
        sc = user.double() @ table[s0:s0 + 50_000].double().t()
        ids = torch.arange(s0, min(V, s0 + 50_000), device=DEV)
        ok = (ids[None, :] != 0) & (ids[None, :] != target[:, None])
        ref += ((sc > ts) & ok).sum(1)
    assert int((ref - c.long()).abs().max()) <= 2           # fp32 vs fp64 near-ties
    assert float((ref == c.long()).float().mean()) > 0.9
