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
