"""Loader for tests/golden/*.npz (written by oracle/make_golden.py from the reference classes)."""
import glob
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
        self.name = name
        self.cfg = json.loads(bytes(z['config_json']).decode())
        self.params = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('param/')}
        self.grads = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('grad/')}
        self.traj_params = {k[11:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('traj_param/')}
        self.batch = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('batch/')}
        self.loss = torch.from_numpy(z['loss'])
        self.loss_vec = torch.from_numpy(z['loss_vec'])
        self.scores = torch.from_numpy(z['scores'])
        self.user_emb = torch.from_numpy(z['user_emb'])
        self.traj_loss = z['traj_loss']
        self.model = self.cfg['model']
        # dropout fixtures: the reference ran with explicit masks of (drop_seed, drop_step + iteration), see oracle/make_golden.py
        self.drop_seed, self.drop_step = self.cfg.get('drop_seed'), self.cfg.get('drop_step')

    def drop_masks(self, it=0):
        """Explicit dropout multipliers of training iteration `it` (None for the dropout-free fixtures)."""
        if self.drop_step is None:
            return None
        from oracle import philox
        B, L = self.batch['item_seq'].shape
        fn = philox.sasrec_masks if self.model == 'SASRec' else philox.gru_masks
        return fn(self.cfg, B, L, self.drop_seed, self.drop_step + it)

    def arm(self, model, it=0):
        """Make the next training forward of a unirec_b200 model draw this fixture's mask set."""
        if self.drop_step is not None:
            model._engine.set_dropout_state(self.drop_seed, self.drop_step + it)

    def fwd_batch(self):
        if self.model == 'MF':
            return {k: self.batch[k] for k in ('user_id', 'item_id', 'label')}
        return dict(self.batch)


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
