"""BaseDataset (drop-in for unirec/data/dataset/basedataset.py for the formats the hot path uses: T1 user-item, T1_1
user-item-max_len, T2 user-item-label).  Data is held COLUMNAR (numpy int64 columns) instead of an object ndarray of rows, so the
same arrays feed both this per-sample CPU view (evaluation, tests) and the device-side batch builder (training)."""
import logging
import os
import pickle

import numpy as np
import pandas as pd
from torch.utils.data import Dataset

from unirec_b200.constants.protocols import DataFileFormat, EvaluationProtocal

_COLUMNS = {
    DataFileFormat.T1.value: ['user_id', 'item_id'],
    DataFileFormat.T1_1.value: ['user_id', 'item_id', 'max_len'],
    DataFileFormat.T2.value: ['user_id', 'item_id', 'label'],
    DataFileFormat.T3.value: ['user_id', 'item_id'],
}


def load_frame(path, filename):
    """<filename>.ftr | .pkl (pickled DataFrame) | .tsv/.csv/.txt, as the reference (basedataset.py:209-226)."""
    base = os.path.join(path, filename)
    if os.path.exists(base + '.ftr'):
        return pd.read_feather(base + '.ftr').reset_index(drop=True)
    if os.path.exists(base + '.pkl'):
        with open(base + '.pkl', 'rb') as f:
            return pickle.load(f).reset_index(drop=True)
    for ext, sep in (('.tsv', '\t'), ('.csv', ','), ('.txt', '\t')):
        if os.path.exists(base + ext):
            return pd.read_csv(base + ext, sep=sep).reset_index(drop=True)
    raise NotImplementedError('no data file %s.{ftr,pkl,tsv,csv,txt}' % base)


class BaseDataset(Dataset):
    def __init__(self, config, path, filename, transform=None):
        self.config = config
        self.logger = logging.getLogger(config['exp_name'])
        fmt = config['data_format']
        if fmt not in _COLUMNS:
            raise ValueError('data format %r is outside the accelerated path (supported: %s)' % (fmt, sorted(_COLUMNS)))
        if config.get('use_features', 0):
            raise ValueError('item features are outside the accelerated path')
        df = load_frame(path, filename)[_COLUMNS[fmt]]
        if config.get('eval_protocol') in (EvaluationProtocal.OneVSAll.value, EvaluationProtocal.OneVSK.value) and 'label' in df:
            df = df[df['label'] > 0]
        self.columns = {c: df[c].to_numpy() for c in df.columns}
        self.user_id = self.columns['user_id'].astype(np.int64)
        self.item_id = self.columns['item_id'].astype(np.int64)
        self.label = self.columns['label'].astype(np.int32) if 'label' in self.columns else None
        self.max_len = self.columns['max_len'].astype(np.int64) if 'max_len' in self.columns else None
        self.transform = transform
        self.set_return_column_index()

    def set_return_column_index(self):
        self.return_key_2_index = {'user_id': 0, 'item_id': 1, 'label': 2}
        if self.config['data_format'] == DataFileFormat.T1_1.value:
            self.return_key_2_index['max_len'] = len(self.return_key_2_index)

    def __len__(self):
        return len(self.user_id)

    def _row(self, index):
        row = [int(self.user_id[index]), int(self.item_id[index])]
        if self.label is not None:
            row.append(int(self.label[index]))
        elif self.max_len is not None:
            row.append(int(self.max_len[index]))
        return np.asarray(row, dtype=object)

    def __getitem__(self, index):
        sample = self._row(index)
        max_len = sample[2] if self.max_len is not None else None
        if self.transform is not None:
            sample = self.transform(sample)
        user_id, item_id = sample[0], sample[1]
        if self.label is not None:
            label = sample[2]
        else:   # fake label: the first candidate is the positive (reference basedataset.py:182-190)
            if isinstance(item_id, np.ndarray):
                label = np.zeros(len(item_id), dtype=np.int32)
                label[0] = 1
            else:
                label = 1
        out = (user_id, item_id, label)
        if max_len is not None:
            out = out + (max_len,)
        return out
