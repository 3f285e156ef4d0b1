"""Data side of the boundary: CPU tests for the dataset / transform contract, GPU tests for the device batch builder and for
`main.run` end to end on a synthetic on-disk dataset in the reference's file format (pickled DataFrames)."""
import os
import pickle

import numpy as np
import pandas as pd
import pytest
import torch

from unirec_b200.data.history import UserHistoryCSR
from unirec_b200.utils import general
from unirec_b200.utils.sampling import build_alias_table, popularity_weights


def make_dataset(root, n_users=60, n_items=200, seed=0):
    """user histories with a planted pattern (user u likes items congruent to u mod 7) in the reference's on-disk format:
    train/valid/test.pkl = DataFrame[user_id, item_id]; user_history.pkl = DataFrame[user_id, item_id] (T1)."""
    rng = np.random.default_rng(seed)
    rows = []
    for u in range(1, n_users):
        pool = np.arange(1, n_items)[np.arange(1, n_items) % 7 == u % 7]
        n = rng.integers(6, 14)
        rows += [(u, int(i)) for i in rng.choice(pool, size=min(n, len(pool)), replace=False)]
    df = pd.DataFrame(rows, columns=['user_id', 'item_id'])
    last = df.groupby('user_id').tail(1)
    prev = df.drop(last.index).groupby('user_id').tail(1)
    train = df.drop(last.index).drop(prev.index)
    os.makedirs(root, exist_ok=True)
    for name, part in (('train', train), ('valid', prev), ('test', last), ('user_history', df)):
        with open(os.path.join(root, name + '.pkl'), 'wb') as f:
            pickle.dump(part.reset_index(drop=True), f)
    return df


def base_config(root, **over):
    cfg = dict(exp_name='t', n_users=60, n_items=200, max_seq_len=8, data_format='user-item', data_loader_task='train',
               eval_protocol=None, use_features=0, time_seq=0, history_mask_mode='autoregressive', dataset_path=root)
    cfg.update(over)
    return cfg


def test_alias_table_reproduces_distribution():
    w = popularity_weights(np.array([5, 1, 2, 3, 0, 9], dtype=float), 1.0)
    prob, alias = build_alias_table(w)
    n = len(w)
    implied = np.zeros(n)
    for s in range(n):
        implied[s] += prob[s] / n
        implied[alias[s]] += (1 - prob[s]) / n
    assert np.allclose(implied, w, atol=1e-6) and implied[0] == 0


def test_csr_roundtrip_and_sorted():
    obj = np.empty(5, dtype=object)
    obj[1], obj[3] = np.array([9, 3, 7]), np.array([2])
    csr = UserHistoryCSR.from_object_array(obj)
    assert csr.ptr.tolist() == [0, 0, 3, 3, 4, 4] and csr.history(1).tolist() == [9, 3, 7]
    assert csr.sorted_items.tolist() == [3, 7, 9, 2]
    back = csr.to_object_array()
    assert back[1].tolist() == [9, 3, 7] and back[0] is None
    csr2 = UserHistoryCSR.from_interactions([3, 1, 1, 1], [2, 9, 3, 7], 5)
    assert csr2.history(1).tolist() == [9, 3, 7] and csr2.history(3).tolist() == [2]


def test_cpu_dataset_contract(tmp_path):
    root = str(tmp_path)
    make_dataset(root)
    hist, _ = general.load_user_history(root, 'user_history', 60, 'user-item')
    cfg = base_config(root)
    AddNeg = general.get_class_instance('AddNegSamples', 'unirec_b200/data')
    AddHist = general.get_class_instance('AddUserHistory', 'unirec_b200/data')
    DS = general.get_class_instance('SeqRecDataset', 'unirec_b200/data/dataset')
    ds = DS(cfg, root, 'train', transform=AddNeg(60, 200, 4, user2history=hist))
    ds.add_user_history_transform(AddHist(hist, 'autoregressive', None, 1, 'user-item'))
    assert ds.return_key_2_index == {'user_id': 0, 'item_id': 1, 'label': 2, 'item_seq': 3, 'item_seq_len': 4}
    batch = next(iter(torch.utils.data.DataLoader(ds, batch_size=5)))
    user_id, item_id, label, item_seq, seq_len = batch
    assert user_id.dtype == torch.int64 and item_id.dtype == torch.int64 and item_id.shape == (5, 5)
    assert label.dtype == torch.int32 and label[:, 0].tolist() == [1] * 5 and int(label[:, 1:].sum()) == 0
    assert item_seq.dtype == torch.int32 and item_seq.shape == (5, 8) and seq_len.dtype == torch.int64
    for b in range(5):
        h = set(hist[int(user_id[b])].tolist())
        assert all(int(x) not in h and int(x) != int(item_id[b, 0]) and 1 <= int(x) < 200 for x in item_id[b, 1:])
        n = int(seq_len[b])
        full = hist[int(user_id[b])].tolist()
        cut = len(full) - 1 - full[::-1].index(int(item_id[b, 0]))        # autoregressive, last occurrence
        assert item_seq[b].tolist() == [0] * (8 - n) + full[:cut][-n:] if n else True


@pytest.mark.gpu
def test_device_batch_builder_matches_cpu_rules(tmp_path):
    from unirec_b200 import ops
    root = str(tmp_path)
    make_dataset(root)
    hist, _ = general.load_user_history(root, 'user_history', 60, 'user-item')
    csr = UserHistoryCSR.from_object_array(hist, 60)
    ptr, items, srt = csr.device_tensors('cuda')
    with open(os.path.join(root, 'train.pkl'), 'rb') as f:
        train = pickle.load(f)
    u = torch.from_numpy(train['user_id'].to_numpy()).cuda()
    pos = torch.from_numpy(train['item_id'].to_numpy()).cuda()
    AddHist = general.get_class_instance('AddUserHistory', 'unirec_b200/data')
    for mode in ('autoregressive', 'unorder', 'autoagressive'):          # the last one is the reference tests' typo: no masking
        item_id, label, seq, seq_len = ops.build_batch(u, pos, 60, 200, 16, 8, ptr, items, srt, mask_mode=mode, seq_last=1, seed=5, step=3)
        ref_t = AddHist(hist, mode, None, 1, 'user-item')
        assert item_id.shape == (len(u), 17) and torch.equal(item_id[:, 0], pos)
        assert label.dtype == torch.int32 and int(label[:, 0].sum()) == len(u) and int(label[:, 1:].sum()) == 0
        for b in range(len(u)):
            h, n, _ = ref_t((int(u[b]), int(pos[b])))
            want = np.zeros(8, dtype=np.int32)
            k = min(len(h), 8)
            if k:
                want[8 - k:] = h[len(h) - k:]
            assert seq[b].cpu().numpy().tolist() == want.tolist(), (mode, b)
            assert int(seq_len[b]) == min(n, 8)
            hs = set(hist[int(u[b])].tolist())
            negs = item_id[b, 1:].tolist()
            assert all(1 <= x < 200 and x not in hs and x != int(pos[b]) for x in negs)
    # same (seed, step) -> same batch; different step -> different negatives
    a = ops.build_batch(u, pos, 60, 200, 16, 8, ptr, items, srt, seed=5, step=3)[0]
    b2 = ops.build_batch(u, pos, 60, 200, 16, 8, ptr, items, srt, seed=5, step=3)[0]
    c = ops.build_batch(u, pos, 60, 200, 16, 8, ptr, items, srt, seed=5, step=4)[0]
    assert torch.equal(a, b2) and not torch.equal(a, c)


@pytest.mark.gpu
def test_device_sampler_distribution_and_exhaustion():
    from unirec_b200 import ops
    B, V, K = 4096, 50, 64
    u = torch.ones(B, dtype=torch.int64, device='cuda')
    pos = torch.full((B,), 7, dtype=torch.int64, device='cuda')
    item_id, *_ = ops.build_batch(u, pos, 4, V, K, 0, seed=1, step=0)
    negs = item_id[:, 1:].reshape(-1).cpu().numpy()
    counts = np.bincount(negs, minlength=V)
    assert counts[0] == 0 and counts[7] == 0
    expect = negs.size / (V - 2)
    assert np.all(np.abs(counts[[i for i in range(1, V) if i != 7]] - expect) < 6 * np.sqrt(expect))
    # popularity sampling follows the alias table
    w = popularity_weights(np.arange(V, dtype=float) ** 2, 0.5)
    prob, alias = build_alias_table(w)
    item_id, *_ = ops.build_batch(u, pos, 4, V, K, 0, alias_prob=torch.from_numpy(prob).cuda(), alias_idx=torch.from_numpy(alias).cuda(),
                                  seed=2, step=0)
    negs = item_id[:, 1:].reshape(-1).cpu().numpy()
    freq = np.bincount(negs, minlength=V) / negs.size
    w2 = w.copy(); w2[7] = 0; w2 /= w2.sum()
    assert np.abs(freq - w2).max() < 0.01
    # every item is either the positive or in the history: all draws fail -> id 0 (reference: addnegsamples.py:99-107)
    ptr = torch.tensor([0, 0, 3, 3, 3], dtype=torch.int64, device='cuda')
    items = torch.tensor([1, 2, 3], dtype=torch.int32, device='cuda')
    pos4 = torch.full((8,), 4, dtype=torch.int64, device='cuda')
    item_id, *_ = ops.build_batch(u[:8], pos4, 4, 5, 6, 0, ptr, items, items, seed=3, step=0)
    assert int(item_id[:, 1:].abs().sum()) == 0


@pytest.mark.gpu
def test_main_run_end_to_end(tmp_path):
    """`main.run` on an on-disk dataset in the reference's format: device batch builder -> SASRec on the CUDA path -> evaluation."""
    from unirec_b200.main import main as entry
    root = str(tmp_path / 'data')
    make_dataset(root)
    res = entry.run(dict(model='SASRec', dataset='example', dataloader='SeqRecDataset', exp_name='e2e', dataset_path=root,
                         output_path=str(tmp_path / 'out'), n_users=60, n_items=200, embedding_size=32, hidden_size=32, n_layers=1,
                         n_heads=2, inner_size=64, max_seq_len=8, hidden_dropout_prob=0.0, attn_dropout_prob=0.0,
                         loss_type='softmax', n_sample_neg_train=20, n_sample_neg_valid=19, n_sample_neg_test=19,
                         train_file_format='user-item', valid_file_format='user-item', test_file_format='user-item',
                         user_history_filename='user_history', user_history_file_format='user-item',
                         history_mask_mode='autoregressive', seq_last=1, epochs=30, batch_size=64, learning_rate=0.01,
                         scheduler='none', early_stop=0, num_workers=0, valid_protocol='one_vs_k', test_protocol='one_vs_k',
                         metrics="['hit@5', 'ndcg@5', 'group_auc']", key_metric='group_auc', gpu_id=-1))
    assert res['group_auc'] > 0.75, res            # the planted (user mod 7 == item mod 7) pattern is learnable
    assert os.path.exists(str(tmp_path / 'out' / 'result_e2e.tsv'))


@pytest.mark.parametrize('tag,kw', [('unorder', dict(mask_mode='unorder')), ('auto_last', dict(mask_mode='autoregressive', seq_last=1)),
                                    ('maxlen', dict(mask_mode='autoregressive', data_format='user-item-max_len'))])
def test_history_transform_matches_reference_golden(tag, kw):
    """a15: AddUserHistory + the dataset's left-padding reproduce the reference transform (fixture written by
    oracle/make_data_golden.py from unirec/data/transform/adduserhistory.py:32-73) on 200 samples, including users without history,
    user ids beyond the table, duplicated history items and targets that occur several times in the history."""
    from unirec_b200.data.dataset.seqrecdataset import SeqRecDataset
    from unirec_b200.data.transform.adduserhistory import AddUserHistory
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden_eval', 'history_transform.npz'))
    n_users, L = int(z['n_users']), int(z['L'])
    hist = np.empty(n_users, dtype=object)
    for u in range(n_users):
        h = z['hist/%d' % u]
        hist[u] = h if len(h) else None
    tr = AddUserHistory(hist, **kw)
    ds = SeqRecDataset.__new__(SeqRecDataset)
    ds.config = {'max_seq_len': L}
    for i, u in enumerate(z['users']):
        sample = (int(u), int(z['targets'][i]), int(z['maxlens'][i])) if tag == 'maxlen' else (int(u), int(z['targets'][i]))
        h, n, _ = tr(sample)
        assert np.array_equal(ds._padding(np.asarray(h)), z['seq/' + tag][i]), (tag, i, h, z['seq/' + tag][i])
        assert min(int(n), L) == int(z['len/' + tag][i]), (tag, i)
