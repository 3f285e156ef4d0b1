"""Generate tests/golden/*.npz from the UNMODIFIED reference classes (TEST INFRASTRUCTURE ONLY).

Run in the build container, where /root/reference exists:

    python oracle/make_golden.py

Each fixture holds: the config (json), a seeded synthetic batch following the batch contract
(user_id i64 [B], item_id i64 [B,1+K] positive first, label i32 [B,1+K], item_seq i32
[B,L] left-padded with 0, item_seq_len i64 [B]), the model's state_dict as initialised by the
reference constructors under `general.init_seed(seed)`, and the reference outputs: loss (mean and
per-sample), scores, user_emb, dense parameter gradients, and a 3-step dense-Adam trajectory
(losses + final parameters) following unirec/facility/trainer.py:340-349.

The GPU box has no /root/reference; only the .npz files travel.
"""
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get('UNIREC_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')

BASE = dict(dataset='example', exp_name='golden', train_file_format='user-item',
            hidden_dropout_prob=0.0, attn_dropout_prob=0.0, dropout_prob=0.0, scheduler='none')

CASES = {
    'sasrec_softmax': dict(model='SASRec', n_items=211, n_users=37, embedding_size=32, hidden_size=32,
                           n_layers=2, n_heads=2, inner_size=64, max_seq_len=8, loss_type='softmax',
                           K=5, B=6, init_std=0.2),
    'sasrec_bpr_nopos_bias': dict(model='SASRec', n_items=157, n_users=29, embedding_size=64, hidden_size=64,
                                  n_layers=1, n_heads=4, inner_size=96, max_seq_len=11, loss_type='bpr',
                                  use_position_emb=0, hidden_act='gelu', tau=0.7, has_user_bias=True,
                                  has_item_bias=True, score_clip_value=0.05, K=3, B=5),
    'sasrec_softmax_d128': dict(model='SASRec', n_items=301, n_users=11, embedding_size=128, hidden_size=128,
                                n_layers=1, n_heads=2, inner_size=256, max_seq_len=50, loss_type='softmax',
                                K=16, B=4, hidden_act='swish', init_std=0.1),
    # padding edge cases of the data path (adduserhistory.py:45-51 'unorder' masking zeroes history items in place, :60-68 can cut
    # the history to length 0): a sequence with no real item, a pad at the output position L-1, pads between real items
    'sasrec_softmax_padedges': dict(model='SASRec', n_items=211, n_users=37, embedding_size=32, hidden_size=32,
                                    n_layers=2, n_heads=2, inner_size=64, max_seq_len=8, loss_type='softmax',
                                    K=5, B=8, init_std=0.2, edge=1),
    'sasrec_bpr_nopos_padedges': dict(model='SASRec', n_items=157, n_users=29, embedding_size=64, hidden_size=64,
                                      n_layers=2, n_heads=4, inner_size=96, max_seq_len=11, loss_type='bpr',
                                      use_position_emb=0, hidden_act='gelu', K=3, B=8, init_std=0.2, edge=1),
    # stock config/model/SASRec.yaml head count (n_heads 16 -> head dim 4 at d = 64)
    'sasrec_h16_d64': dict(model='SASRec', n_items=181, n_users=23, embedding_size=64, hidden_size=64, n_layers=2, n_heads=16,
                           inner_size=128, max_seq_len=12, loss_type='softmax', K=6, B=5, init_std=0.2),
    # dropout > 0: the reference's nn.Dropout modules are replaced by explicit multipliers computed by oracle/philox.py (the
    # generator the CUDA kernels implement), so loss / grads / trajectory are the reference's arithmetic on a known mask set
    'sasrec_softmax_drop': dict(model='SASRec', n_items=211, n_users=37, embedding_size=32, hidden_size=32, n_layers=2, n_heads=2,
                                inner_size=64, max_seq_len=8, loss_type='softmax', K=5, B=8, init_std=0.2, edge=1,
                                hidden_dropout_prob=0.3, attn_dropout_prob=0.2, drop_step=5),
    # the stock SASRec.yaml recipe: 16 heads (head dim 2 at d = 32), dropout 0.5 / 0.5, swish
    'sasrec_stock_drop': dict(model='SASRec', n_items=157, n_users=29, embedding_size=32, hidden_size=32, n_layers=2, n_heads=16,
                              inner_size=64, max_seq_len=10, loss_type='softmax', K=4, B=6, init_std=0.2,
                              hidden_dropout_prob=0.5, attn_dropout_prob=0.5, drop_step=11),
    'gru_bpr_drop': dict(model='GRU', n_items=123, n_users=19, embedding_size=32, hidden_size=64, max_seq_len=7, loss_type='bpr',
                         K=5, B=6, init_std=0.3, dropout_prob=0.4, drop_step=3),
    # BASELINE config c3 shape of the recurrence (d = h = 256, L = 100) on a small catalogue: pins the persistent GRU kernel
    'gru_d256_h256_L100': dict(model='GRU', n_items=301, n_users=19, embedding_size=256, hidden_size=256, max_seq_len=100,
                               loss_type='bpr', K=5, B=3, init_std=0.1),
    'gru_bpr': dict(model='GRU', n_items=123, n_users=19, embedding_size=32, hidden_size=64,
                    max_seq_len=7, loss_type='bpr', K=5, B=6, init_std=0.3),
    'gru_softmax_h32': dict(model='GRU', n_items=99, n_users=19, embedding_size=32, hidden_size=32,
                            max_seq_len=5, loss_type='softmax', K=4, B=3, tau=0.5),
    'avghist_softmax': dict(model='AvgHist', n_items=173, n_users=23, embedding_size=32,
                            max_seq_len=12, loss_type='softmax', K=7, B=6, init_std=0.3),
    'avghist_sym_bpr': dict(model='AvgHist', n_items=173, n_users=23, embedding_size=64, asymmetric=False,
                            max_seq_len=9, loss_type='bpr', K=2, B=5),
    'svdpp_bpr': dict(model='SVDPlusPlus', n_items=143, n_users=31, embedding_size=32,
                      max_seq_len=10, loss_type='bpr', K=5, B=7),
    'mf_bpr': dict(model='MF', n_items=97, n_users=41, embedding_size=64, loss_type='bpr', K=1, B=16, init_std=0.3),
    'mf_softmax_bias': dict(model='MF', n_items=97, n_users=41, embedding_size=32, loss_type='softmax',
                            has_item_bias=True, K=6, B=9, tau=2.0),
}


def make_batch(cfg, B, K, L, gen, edge=False):
    V, U = cfg['n_items'], cfg['n_users']
    user_id = torch.randint(1, U, (B,), generator=gen, dtype=torch.int64)
    item_id = torch.randint(1, V, (B, 1 + K), generator=gen, dtype=torch.int64)
    # edge cases the reference data path produces: a failed negative draw yields id 0
    # (addnegsamples.py:99-107); duplicated negatives; a target that also sits in the history.
    if K >= 2:
        item_id[0, K] = 0
        item_id[1, 2] = item_id[1, 1]
    label = torch.zeros(B, 1 + K, dtype=torch.int32)
    label[:, 0] = 1
    item_seq = torch.zeros(B, L, dtype=torch.int32)
    lens = torch.randint(1, L + 1, (B,), generator=gen, dtype=torch.int64)
    lens[0] = L
    if B > 1:
        lens[1] = 1
    for b in range(B):
        n = int(lens[b])
        item_seq[b, L - n:] = torch.randint(1, V, (n,), generator=gen, dtype=torch.int64).to(torch.int32)
    if B > 2 and L > 1:
        item_seq[2, L - 1] = item_id[2, 0].to(torch.int32)
    if edge:
        item_seq[3, :] = 0                 # empty history
        lens[3] = 0
        item_seq[4, L - 4:] = torch.randint(1, V, (4,), generator=gen, dtype=torch.int64).to(torch.int32)
        item_seq[4, L - 1] = 0             # pad at the output position, real items before it
        lens[4] = 4
        item_seq[5, L - 5:] = torch.randint(1, V, (5,), generator=gen, dtype=torch.int64).to(torch.int32)
        item_seq[5, L - 3] = 0             # pad between real items
        lens[5] = 5
    return dict(user_id=user_id, item_id=item_id, label=label, item_seq=item_seq, item_seq_len=lens)


class InjectedDropout(torch.nn.Module):
    """Stands in for one nn.Dropout of the reference model: multiplies by the explicit mask `store[key]`."""

    def __init__(self, store, key):
        super().__init__()
        self.store, self.key = store, key

    def forward(self, x):
        return x * self.store[self.key].view(x.shape) if self.training else x


def inject_dropout(model, name, store):
    """Replace every nn.Dropout on the tower by an InjectedDropout reading `store` (keys of oracle/philox.py)."""
    if name == 'SASRec':
        model.dropout = InjectedDropout(store, 'input')                       # sasrec.py:69
        for i, layer in enumerate(model.trm_encoder.layer):
            layer.multi_head_attention.attn_dropout = InjectedDropout(store, 'attn.%d' % i)       # modules.py:307
            layer.multi_head_attention.out_dropout = InjectedDropout(store, 'attn_out.%d' % i)    # modules.py:313
            layer.feed_forward.dropout = InjectedDropout(store, 'ffn_out.%d' % i)                 # modules.py:352
    elif name == 'GRU':
        model.emb_dropout = InjectedDropout(store, 'input')                   # gru.py:29
    else:
        raise ValueError(name)
    assert not any(isinstance(m, torch.nn.Dropout) and m.p > 0 for m in model.modules())


def main():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
    from oracle import philox
    from unirec.utils import argument_parser, general
    os.makedirs(OUT, exist_ok=True)
    saved_argv, sys.argv = sys.argv, sys.argv[:1]
    only = set(saved_argv[1:])
    for name, case in CASES.items():
        if only and name not in only:
            continue
        case = dict(case)
        B, K = case.pop('B'), case.pop('K')
        edge = bool(case.pop('edge', 0))
        drop_step = case.pop('drop_step', None)
        args = dict(BASE)
        args.update(case)
        cfg = argument_parser.parse_arguments(args)
        cfg['device'] = torch.device('cpu')
        seed = 2022
        general.init_seed(seed)
        model = general.get_class_instance(cfg['model'], 'unirec/model')(cfg)
        model.train()
        L = int(cfg['max_seq_len'])
        gen = torch.Generator().manual_seed(seed + 1)
        batch = make_batch(cfg, B, K, L, gen, edge)
        if cfg['model'] == 'MF':
            fwd_batch = {k: batch[k] for k in ('user_id', 'item_id', 'label')}
        else:
            fwd_batch = batch
        out = {}
        for k, v in model.state_dict().items():
            out['param/' + k] = v.detach().numpy().copy()
        store = {}

        def set_masks(step):
            if drop_step is None:
                return
            fn = philox.sasrec_masks if cfg['model'] == 'SASRec' else philox.gru_masks
            store.clear()
            store.update(fn(cfg, B, L, seed, step))

        if drop_step is not None:
            inject_dropout(model, cfg['model'], store)
            set_masks(drop_step)
        loss, scores, user_emb, items_emb = model(**fwd_batch, return_loss_only=False)
        model.zero_grad()
        loss.backward()
        for k, p_ in model.named_parameters():
            g = p_.grad if p_.grad is not None else torch.zeros_like(p_)
            out['grad/' + k] = g.detach().numpy().copy()
        loss_vec = model(**fwd_batch, reduction=False)[0]
        out['loss'] = loss.detach().numpy()
        out['loss_vec'] = loss_vec.detach().numpy()
        out['scores'] = scores.detach().numpy()
        out['user_emb'] = user_emb.detach().numpy()
        # 3-step dense Adam trajectory on the same batch (trainer.py:136,340-349)
        opt = torch.optim.Adam(model.parameters(), lr=float(cfg['learning_rate']),
                               weight_decay=float(cfg['weight_decay']))
        traj = []
        for it in range(3):
            set_masks((drop_step or 0) + it)            # one mask set per training step
            l_ = model(**fwd_batch)[0]
            opt.zero_grad()
            l_.backward()
            opt.step()
            traj.append(float(l_.detach()))
        out['traj_loss'] = np.asarray(traj, dtype=np.float64)
        for k, v in model.state_dict().items():
            out['traj_param/' + k] = v.detach().numpy().copy()
        for k, v in batch.items():
            out['batch/' + k] = v.numpy()
        keep = {k: v for k, v in cfg.items()
                if isinstance(v, (int, float, str, bool)) and k not in ('exp_name', 'config_dir')}
        keep['K'], keep['B'] = K, B
        if drop_step is not None:
            keep['drop_seed'], keep['drop_step'] = seed, drop_step
        out['config_json'] = np.frombuffer(json.dumps(keep, sort_keys=True).encode(), dtype=np.uint8)
        path = os.path.join(OUT, name + '.npz')
        np.savez_compressed(path, **out)
        print('%-26s loss=%.6f  %7.1f KB' % (name, float(loss), os.path.getsize(path) / 1024))
    sys.argv = saved_argv


if __name__ == '__main__':
    main()
