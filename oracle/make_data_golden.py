"""Generate tests/golden_eval/history_transform.npz from the UNMODIFIED reference data transforms (TEST INFRASTRUCTURE ONLY).

    python oracle/make_data_golden.py        # in the build container, where /root/reference exists

Deterministic parts of the batch producers: `AddUserHistory.__call__` (unirec/data/transform/adduserhistory.py:32-73)
in its three deterministic modes (unorder; autoregressive with seq_last=1; user-item-max_len) and the left-padding of
`SeqRecDataset._padding` (unirec/data/dataset/seqrecdataset.py:60-68, restated here because the dataset module imports `feather`,
which this image lacks -- the function body is the 7 lines cited).
"""
import os
import sys

import numpy as np

REF = os.environ.get('UNIREC_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden_eval')


def ref_padding(x, k):
    res = np.zeros((k,), dtype=np.int32)
    if len(x) < k:
        res[(k - len(x)):] = x[:]
    else:
        res[:] = x[len(x) - k:]
    return res


def main():
    sys.path.insert(0, REF)
    from unirec.data.transform.adduserhistory import AddUserHistory
    rng = np.random.RandomState(11)
    n_users, n_items, L = 40, 60, 8
    hist = np.empty(n_users, dtype=object)
    for u in range(n_users):
        if u in (0, 5):                         # user 0 is padding; user 5 has no history
            hist[u] = None
            continue
        n = int(rng.randint(1, 15))
        hist[u] = rng.randint(1, n_items, size=n).astype(np.int32)      # duplicates allowed
    users = rng.randint(1, n_users + 2, size=200)                       # includes ids beyond the table (no history)
    out = {'n_users': np.int64(n_users), 'L': np.int64(L), 'users': users}
    for u in range(n_users):
        out['hist/%d' % u] = np.zeros(0, np.int32) if hist[u] is None else hist[u]
    # targets: mostly an item of the user's own history (so masking/cutting has something to do)
    targets = np.zeros(len(users), dtype=np.int64)
    maxlens = np.zeros(len(users), dtype=np.int64)
    for i, u in enumerate(users):
        h = hist[u] if u < n_users and hist[u] is not None else np.zeros(0, np.int32)
        targets[i] = int(h[rng.randint(len(h))]) if len(h) and rng.rand() < 0.8 else int(rng.randint(1, n_items))
        maxlens[i] = int(rng.randint(0, len(h) + 1)) if len(h) else 0
    out['targets'], out['maxlens'] = targets, maxlens
    modes = {'unorder': dict(mask_mode='unorder'), 'auto_last': dict(mask_mode='autoregressive', seq_last=1),
             'maxlen': dict(mask_mode='autoregressive', data_format='user-item-max_len')}
    for tag, kw in modes.items():
        tr = AddUserHistory(hist, **kw)
        seqs, lens = [], []
        for i, u in enumerate(users):
            sample = (int(u), int(targets[i]), int(maxlens[i])) if tag == 'maxlen' else (int(u), int(targets[i]))
            h, n, _ = tr(sample)
            seqs.append(ref_padding(np.asarray(h), L))
            lens.append(min(int(n), L))
        out['seq/' + tag] = np.stack(seqs)
        out['len/' + tag] = np.asarray(lens, dtype=np.int64)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, 'history_transform.npz')
    np.savez_compressed(path, **out)
    print('wrote', path)


if __name__ == '__main__':
    main()
