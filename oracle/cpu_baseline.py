"""CPU baseline leg (TEST / MEASUREMENT INFRASTRUCTURE ONLY): the reference's training step, restated by the
oracle, timed on the host cores.  Used by bench.py (`cpu_baseline` object and `--impl reference`).

Step timed = unirec/facility/trainer.py:340-349 with torch.optim.Adam over ALL parameters (dense table gradients,
dense Adam), i.e. `loss = model(**batch)[0]; opt.zero_grad(); loss.backward(); opt.step()`; forward arithmetic is
oracle.unirec_oracle.forward.  /root/reference does not exist on the GPU box, so the reference classes themselves
cannot be timed there; this port was pinned against them by tests/golden (kind = "port").
"""
import os
import time

import torch

from . import unirec_oracle as O


def init_params(model, cfg, seed=2022):
    """Random-init parameters of the named architecture (N(0, init_std) like the reference, padding rows zero)."""
    g = torch.Generator().manual_seed(seed)
    V, U, d = int(cfg['n_items']), int(cfg['n_users']), int(cfg['embedding_size'])
    std = float(cfg.get('init_std', 0.02))

    def normal(*shape):
        return torch.empty(*shape).normal_(0.0, std, generator=g)

    def table(n):
        t = normal(n, d)
        t[0] = 0
        return t

    p = {'item_embedding.weight': table(V)}
    if model in ('MF', 'SVDPlusPlus'):
        p['user_embedding.weight'] = table(U)
    if model == 'SVDPlusPlus' or (model == 'AvgHist' and cfg.get('asymmetric', True)):
        p['item_dst_embedding.weight'] = table(V)
    if model == 'SASRec':
        L, I = int(cfg['max_seq_len']), int(cfg['inner_size'])
        if cfg.get('use_position_emb', True):
            p['position_embedding.weight'] = normal(L + 1, d)
        p['LayerNorm.weight'], p['LayerNorm.bias'] = torch.ones(d), torch.zeros(d)
        for i in range(int(cfg['n_layers'])):
            a, f = 'trm_encoder.layer.%d.multi_head_attention.' % i, 'trm_encoder.layer.%d.feed_forward.' % i
            for n in ('query', 'key', 'value', 'dense'):
                p[a + n + '.weight'], p[a + n + '.bias'] = normal(d, d), torch.zeros(d)
            p[a + 'LayerNorm.weight'], p[a + 'LayerNorm.bias'] = torch.ones(d), torch.zeros(d)
            p[f + 'dense_1.weight'], p[f + 'dense_1.bias'] = normal(I, d), torch.zeros(I)
            p[f + 'dense_2.weight'], p[f + 'dense_2.bias'] = normal(d, I), torch.zeros(d)
            p[f + 'LayerNorm.weight'], p[f + 'LayerNorm.bias'] = torch.ones(d), torch.zeros(d)
    if model == 'GRU':
        H = int(cfg.get('hidden_size', d))
        k = 1.0 / H ** 0.5
        for n, shape in (('weight_ih_l0', (3 * H, d)), ('weight_hh_l0', (3 * H, H)), ('bias_ih_l0', (3 * H,)), ('bias_hh_l0', (3 * H,))):
            p['gru_layers.' + n] = torch.empty(*shape).uniform_(-k, k, generator=g)
        p['dense.weight'], p['dense.bias'] = normal(d, H), torch.zeros(d)
    return p


def time_steps(model, cfg, batches, steps, warmup, threads=None, lr=1e-3):
    """Returns (samples_per_s, ms_per_step, threads).  `batches`: list of CPU batch dicts, cycled."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    p = init_params(model, cfg)
    leaves = {k: v.requires_grad_(True) for k, v in p.items()}
    opt = torch.optim.Adam(list(leaves.values()), lr=lr, weight_decay=float(cfg.get('weight_decay', 0.0)))
    B = next(iter(batches[0].values())).shape[0]

    def one(i):
        batch = batches[i % len(batches)]
        loss = O.forward(model, leaves, cfg, **batch)[0]
        opt.zero_grad()
        loss.backward()
        with torch.no_grad():
            for k in O.PADDING_TABLES:
                if k in leaves and leaves[k].grad is not None:
                    leaves[k].grad[0] = 0          # nn.Embedding(padding_idx=0)
        opt.step()
        return float(loss.detach())

    for i in range(warmup):
        one(i)
    t0 = time.perf_counter()
    for i in range(steps):
        one(warmup + i)
    dt = time.perf_counter() - t0
    return B * steps / dt, 1e3 * dt / steps, threads
