"""Registry, seeding and checkpoint helpers whose behaviour is API (reference: unirec/utils/general.py)."""
import importlib
import importlib.util
import os
import random
import time

import numpy as np
import torch


def get_local_time_str():
    return time.strftime('%Y-%m-%d-%H:%M:%S', time.localtime())


def dict2str(d, sep=' '):
    return sep.join('{0}:{1}'.format(k, d[k]) for k in sorted(d))


def init_seed(seed):
    """Seed python / numpy / torch (CPU and every CUDA device).  reference: general.py:26-39."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.deterministic = True


_ALIASES = {'unirec': 'unirec_b200'}


def get_class_instance(class_name, class_root='model'):
    """Filename-based registry (reference: general.py:74-103): walk the package `class_root`
    ('unirec_b200/model', or the reference spelling 'unirec/model'), import the first module named
    `class_name.lower()` and return its attribute `class_name`; ValueError when absent."""
    root = class_root.replace('/', '.').replace('\\', '.')
    head, _, tail = root.partition('.')
    root = _ALIASES.get(head, head) + ('.' + tail if tail else '')
    base = os.path.dirname(importlib.import_module(root).__file__)
    wanted = class_name.lower()
    for cur, _dirs, _files in os.walk(base, topdown=True):
        rel = os.path.relpath(cur, base)
        mod = root if rel == '.' else root + '.' + rel.replace(os.sep, '.')
        path = mod + '.' + wanted
        try:
            found = importlib.util.find_spec(path) is not None
        except (ImportError, ValueError):
            found = False
        if found:
            return getattr(importlib.import_module(path), class_name)
    raise ValueError('Cannot import `class name` [{0}] from {1}.'.format(class_name, root))


def load_model_freely(filename, device=None):
    """Rebuild a model from the config stored in its checkpoint and load the weights (reference: general.py:208-230)."""
    cpt = torch.load(filename, map_location=device, weights_only=False) if device is not None \
        else torch.load(filename, weights_only=False)
    cfg = cpt['config']
    if device is None:
        device = cfg['device']
    else:
        cfg['device'] = device
    for k in ('item_emb_path', 'text_emb_path'):
        cfg.pop(k, None)
    model = get_class_instance(cfg['model'], 'unirec_b200/model')(cfg).to(device)
    model.load_state_dict(cpt['state_dict'], strict=False)
    return model, cfg


def load_user_history(file_path, file_name, n_users=None, format='user-item', time_seq=0):
    """User histories as the reference returns them (general.py:111-149): an object ndarray indexed by user id whose entries are
    per-user item arrays (None for users without history), plus the time histories (not supported here -> None).
    `unirec_b200.data.history.UserHistoryCSR.from_object_array` turns it into the CSR form the device path uses."""
    import pandas as pd
    from unirec_b200.constants.protocols import ColNames, DataFileFormat
    from unirec_b200.data.dataset.basedataset import load_frame
    if time_seq:
        raise ValueError('time sequences are outside the accelerated path')
    df = load_frame(file_path, file_name)
    if format in (DataFileFormat.T1.value, DataFileFormat.T3.value):
        grouped = df.groupby('user_id')['item_id'].apply(lambda x: np.array(x))
    elif format in (DataFileFormat.T5.value, DataFileFormat.T6.value):
        grouped = df.set_index('user_id')[ColNames.USER_HISTORY.value].to_dict()
    else:
        raise NotImplementedError('Unsupport user history format: {0}'.format(format))
    if n_users is None or n_users <= 0:
        n_users = int(df['user_id'].max()) + 1
    res = np.empty(n_users, dtype=object)
    for user_id, items in grouped.items():
        res[user_id] = np.asarray(items)
    return res, None
