"""CPU oracle: functional restatement of the UniRec hot path (TEST INFRASTRUCTURE ONLY).

Plain torch-CPU fp32 (or fp64) tensor arithmetic, no nn.Module from the
reference, no CUDA.  Every function cites the reference lines it restates
(paths relative to /root/reference/).  The arithmetic the reference delegates
to PyTorch (embedding, Linear, LayerNorm, softmax, GRU cell, Adam) is restated
from PyTorch's published definitions; parity is pinned by `tests/golden/*.npz`
(see oracle/make_golden.py), which were produced by the reference classes.

Parameters are passed as a flat dict keyed by the reference's `state_dict`
names, so a reference checkpoint feeds the oracle as is.
"""
import math
from typing import Dict, Optional

import torch

EPS = 1e-8  # unirec/constants/global_variables.py:4


# --------------------------------------------------------------------------
# a2 / a3: embedding row gathers
# --------------------------------------------------------------------------
def gather_rows(table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """`nn.Embedding` forward = row gather.  unirec/model/base/recommender.py:66-67,136-137."""
    return table[idx.long()]


# --------------------------------------------------------------------------
# a8 / a9: scorer + bias / temperature / clamp
# --------------------------------------------------------------------------
def inner_product_scores(user_emb: torch.Tensor, items_emb: torch.Tensor) -> torch.Tensor:
    """s[b,j] = <items[b,j,:], u[b,:]>.  unirec/model/modules.py:45-67 (case 3 -> case 2),
    plus the `[B,D]x[B,D]` one-vs-one case (:52-53)."""
    if items_emb.dim() == user_emb.dim():
        return (user_emb * items_emb).sum(-1)
    return torch.matmul(items_emb, user_emb.unsqueeze(-1)).squeeze(-1)


def predict_layer(scores, user_id=None, item_id=None, user_bias=None, item_bias=None,
                  tau: float = 1.0, score_clip: float = -1.0):
    """unirec/model/base/recommender.py:76-96."""
    if user_bias is not None:
        ub = user_bias[user_id.long()]
        if ub.shape != scores.shape:
            ub = ub.unsqueeze(1).expand_as(scores)
        scores = scores + ub
    if item_bias is not None:
        scores = scores + item_bias[item_id.long()]
    scores = scores / tau
    if score_clip > 0:
        scores = torch.clamp(scores, min=-score_clip, max=score_clip)
    return scores


# --------------------------------------------------------------------------
# a10: losses
# --------------------------------------------------------------------------
def softmax_loss(scores, label, reduction=True):
    """(-log_softmax(s))[label>0] (.mean()).  unirec/model/base/reco_abc.py:260-265."""
    nls = -(scores - torch.logsumexp(scores, dim=-1, keepdim=True))
    loss = nls[label > 0]
    return loss.mean() if reduction else loss


def bpr_loss(scores, reduction=True):
    """-log(1e-8 + sigmoid(s0 - sj)).  reco_abc.py:252-255, unirec/model/modules.py:15-21."""
    neg = scores[:, 1:]
    pos = scores[:, 0:1].expand_as(neg)
    loss = -torch.log(EPS + torch.sigmoid(pos - neg))
    return loss.mean() if reduction else loss.mean(dim=-1)


def cal_loss(scores, label, loss_type, reduction=True, group_size=-1):
    """Loss dispatch.  reco_abc.py:220-272 (softmax and bpr branches; the others are out of scope)."""
    if group_size > 0:
        scores = scores.view(-1, group_size)
        if label is not None:
            label = label.view(-1, group_size)
    if loss_type == 'softmax':
        return softmax_loss(scores, label, reduction)
    if loss_type == 'bpr':
        return bpr_loss(scores, reduction)
    raise ValueError('oracle covers loss_type softmax|bpr, got %r' % (loss_type,))


# --------------------------------------------------------------------------
# a4 / a5: SASRec tower
# --------------------------------------------------------------------------
def layer_norm(x, weight, bias, eps):
    """torch.nn.LayerNorm over the last dim (biased variance)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * weight + bias


def linear(x, w, b=None):
    y = torch.matmul(x, w.t())
    return y if b is None else y + b


def activation(x, name):
    """unirec/model/modules.py:337-346."""
    if name == 'swish':
        return x * torch.sigmoid(x)
    if name == 'gelu':
        return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))
    if name == 'relu':
        return torch.relu(x)
    if name == 'tanh':
        return torch.tanh(x)
    if name == 'sigmoid':
        return torch.sigmoid(x)
    raise ValueError(name)


def sasrec_attention_mask(item_seq, causal: bool, dtype):
    """Additive mask [B,1,L,L]: 0 where key is a real item (and key<=query when causal), else -10000.
    unirec/model/sequential/sasrec.py:40-57."""
    key_ok = (item_seq > 0).to(dtype)[:, None, None, :]
    if causal:
        L = item_seq.shape[1]
        tri = torch.tril(torch.ones(L, L, dtype=dtype, device=item_seq.device))[None, None]
        key_ok = key_ok * tri
    return (1.0 - key_ok) * -10000.0


def multi_head_attention(x, mask, p: Dict[str, torch.Tensor], prefix: str, n_heads: int, eps: float, m_attn=None, m_out=None):
    """unirec/model/modules.py:284-316.  m_attn [B,H,L,L] / m_out [B,L,D]: explicit multipliers standing in for
    attn_dropout (:307) and out_dropout (:313); None = identity (eval mode / p = 0)."""
    B, L, D = x.shape
    dh = D // n_heads

    def split(t):
        return t.view(B, L, n_heads, dh).permute(0, 2, 1, 3)

    q = split(linear(x, p[prefix + 'query.weight'], p[prefix + 'query.bias']))
    k = split(linear(x, p[prefix + 'key.weight'], p[prefix + 'key.bias']))
    v = split(linear(x, p[prefix + 'value.weight'], p[prefix + 'value.bias']))
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dh) + mask
    a = torch.softmax(s, dim=-1)
    if m_attn is not None:
        a = a * m_attn
    ctx = torch.matmul(a, v).permute(0, 2, 1, 3).contiguous().view(B, L, D)
    h = linear(ctx, p[prefix + 'dense.weight'], p[prefix + 'dense.bias'])
    if m_out is not None:
        h = h * m_out
    return layer_norm(h + x, p[prefix + 'LayerNorm.weight'], p[prefix + 'LayerNorm.bias'], eps)


def feed_forward(x, p, prefix, act: str, eps: float, m_out=None):
    """unirec/model/modules.py:347-355 (m_out: multipliers standing in for the dropout at :352)."""
    h = activation(linear(x, p[prefix + 'dense_1.weight'], p[prefix + 'dense_1.bias']), act)
    h = linear(h, p[prefix + 'dense_2.weight'], p[prefix + 'dense_2.bias'])
    if m_out is not None:
        h = h * m_out
    return layer_norm(h + x, p[prefix + 'LayerNorm.weight'], p[prefix + 'LayerNorm.bias'], eps)


def sasrec_user_emb(p, cfg, item_seq, drop=None):
    """unirec/model/sequential/sasrec.py:59-76 (+ recommender.py:136-137 for the gather).  drop: dict of explicit dropout
    multipliers (oracle/philox.py:sasrec_masks) or None."""
    drop = drop or {}
    eps = float(cfg['layer_norm_eps'])
    causal = bool(cfg.get('use_position_emb', True))
    x = gather_rows(p['item_embedding.weight'], item_seq)
    if causal:
        L = item_seq.shape[1]
        x = x + p['position_embedding.weight'][:L][None]
    x = layer_norm(x, p['LayerNorm.weight'], p['LayerNorm.bias'], eps)
    if drop.get('input') is not None:                      # sasrec.py:69
        x = x * drop['input']
    mask = sasrec_attention_mask(item_seq, causal, x.dtype)
    for i in range(int(cfg['n_layers'])):
        pre = 'trm_encoder.layer.%d.' % i
        x = multi_head_attention(x, mask, p, pre + 'multi_head_attention.', int(cfg['n_heads']), eps,
                                 drop.get('attn.%d' % i), drop.get('attn_out.%d' % i))
        x = feed_forward(x, p, pre + 'feed_forward.', cfg['hidden_act'], eps, drop.get('ffn_out.%d' % i))
    return x[:, -1, :]


# --------------------------------------------------------------------------
# a6: GRU tower
# --------------------------------------------------------------------------
def gru_user_emb(p, cfg, item_seq, drop=None):
    """unirec/model/sequential/gru.py:27-35.  nn.GRU(batch_first, 1 layer) restated from the PyTorch
    definition: r,z,n gate order; n = tanh(W_in x + b_in + r*(W_hn h + b_hn)); h' = (1-z)*n + z*h.
    drop['input']: explicit multipliers standing in for emb_dropout (gru.py:29)."""
    x = gather_rows(p['item_embedding.weight'], item_seq)
    if drop and drop.get('input') is not None:
        x = x * drop['input']
    w_ih, w_hh = p['gru_layers.weight_ih_l0'], p['gru_layers.weight_hh_l0']
    b_ih, b_hh = p['gru_layers.bias_ih_l0'], p['gru_layers.bias_hh_l0']
    H = w_hh.shape[1]
    B, L, _ = x.shape
    h = torch.zeros(B, H, dtype=x.dtype, device=x.device)
    gi_all = linear(x, w_ih, b_ih)
    for t in range(L):
        gi = gi_all[:, t]
        gh = linear(h, w_hh, b_hh)
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        h = (1.0 - z) * n + z * h
    # the reference applies `dense` to all L outputs and keeps [:, -1]; only the last step is live.
    return linear(h, p['dense.weight'], p['dense.bias'])


# --------------------------------------------------------------------------
# a7: sum-pool towers and MF
# --------------------------------------------------------------------------
def avghist_user_emb(p, cfg, item_seq, item_seq_len):
    """unirec/model/sequential/avghist.py:34-42 (time_seq off)."""
    # asymmetric: the history side reads a separate deep-copied table (avghist.py:16-22); otherwise
    # item_src/item_dst are aliases of item_embedding (all three names appear in the state_dict).
    key = 'item_dst_embedding.weight' if cfg.get('asymmetric', True) else 'item_embedding.weight'
    e = gather_rows(p[key], item_seq)
    coeff = torch.pow((item_seq_len + 1).to(e.dtype), -float(cfg['user_sequence_alpha'])).unsqueeze(1)
    return coeff * e.sum(1)


def svdpp_user_emb(p, cfg, user_id, item_seq, item_seq_len):
    """unirec/model/sequential/svdplusplus.py:31-39."""
    e = gather_rows(p['item_dst_embedding.weight'], item_seq)
    coeff = torch.pow((item_seq_len + 1).to(e.dtype), -float(cfg['user_sequence_alpha'])).unsqueeze(1)
    return gather_rows(p['user_embedding.weight'], user_id) + coeff * e.sum(1)


def mf_user_emb(p, cfg, user_id):
    """unirec/model/base/recommender.py:42-44; unirec/model/cf/mf.py:6-8."""
    return gather_rows(p['user_embedding.weight'], user_id)


# --------------------------------------------------------------------------
# a11: forward orchestration
# --------------------------------------------------------------------------
def forward_user_emb(model: str, p, cfg, user_id=None, item_seq=None, item_seq_len=None, drop=None):
    if model == 'SASRec':
        return sasrec_user_emb(p, cfg, item_seq, drop)
    if model == 'GRU':
        return gru_user_emb(p, cfg, item_seq, drop)
    if model == 'AvgHist':
        return avghist_user_emb(p, cfg, item_seq, item_seq_len)
    if model == 'SVDPlusPlus':
        return svdpp_user_emb(p, cfg, user_id, item_seq, item_seq_len)
    if model == 'MF':
        return mf_user_emb(p, cfg, user_id)
    raise ValueError(model)


def forward(model: str, p, cfg, user_id=None, item_id=None, label=None, item_seq=None,
            item_seq_len=None, reduction=True, drop=None):
    """BaseRecommender.forward (training branch).  unirec/model/base/recommender.py:46-64.
    Returns (loss, scores, user_emb, items_emb)."""
    items_emb = gather_rows(p['item_embedding.weight'], item_id)   # src table for AvgHist/SVD++ too
    user_emb = forward_user_emb(model, p, cfg, user_id, item_seq, item_seq_len, drop)
    scores = inner_product_scores(user_emb, items_emb)
    scores = predict_layer(scores, user_id, item_id,
                           p.get('user_bias') if cfg.get('has_user_bias') else None,
                           p.get('item_bias') if cfg.get('has_item_bias') else None,
                           float(cfg.get('tau', 1.0)), float(cfg.get('score_clip_value', -1) or -1))
    loss = cal_loss(scores, label, cfg['loss_type'], reduction, int(cfg.get('group_size', -1) or -1))
    return loss, scores, user_emb, items_emb


# --------------------------------------------------------------------------
# a12 / a13 / a14: backward + optimizer = one training step
# --------------------------------------------------------------------------
def tie_aliases(model, cfg, p):
    """AvgHist / SVD++ register `item_src_embedding` (and, when symmetric, `item_dst_embedding`) as
    aliases of `item_embedding` (avghist.py:16-22, svdplusplus.py:17-19): one Parameter, several
    state_dict names.  Make the dict entries share storage the same way."""
    if 'item_src_embedding.weight' in p:
        p['item_src_embedding.weight'] = p['item_embedding.weight']
    if model == 'AvgHist' and not cfg.get('asymmetric', True) and 'item_dst_embedding.weight' in p:
        p['item_dst_embedding.weight'] = p['item_embedding.weight']
    return p


PADDING_TABLES = ('item_embedding.weight', 'item_dst_embedding.weight', 'user_embedding.weight')


def loss_and_grads(model, p, cfg, batch, drop=None):
    """Autograd backward of `forward`, with the `padding_idx=0` rule of nn.Embedding: the gradient
    row 0 of every padded table is zero (reco_abc.py:167-170)."""
    leaves = tie_aliases(model, cfg, {k: v.detach().clone().requires_grad_(True) for k, v in p.items()})
    loss, scores, user_emb, _ = forward(model, leaves, cfg, drop=drop, **batch)
    loss.backward()
    grads = {}
    for k, v in leaves.items():
        g = v.grad if v.grad is not None else torch.zeros_like(v)
        if k in PADDING_TABLES:
            g = g.clone()
            g[0] = 0
        grads[k] = g
    return loss.detach(), scores.detach(), user_emb.detach(), grads


class DenseAdam:
    """torch.optim.Adam(lr, betas=(0.9,0.999), eps=1e-8, weight_decay=wd) restated
    (unirec/facility/trainer.py:136): every row of every table moves every step."""

    def __init__(self, params: Dict[str, torch.Tensor], lr=1e-3, weight_decay=0.0,
                 beta1=0.9, beta2=0.999, eps=1e-8):
        self.lr, self.wd, self.b1, self.b2, self.eps = lr, weight_decay, beta1, beta2, eps
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.t = 0

    def step(self, params, grads):
        self.t += 1
        bc1 = 1.0 - self.b1 ** self.t
        bc2 = 1.0 - self.b2 ** self.t
        done = set()
        for k, p in params.items():
            if id(p) in done:      # aliased names of one Parameter are stepped once
                continue
            done.add(id(p))
            g = grads[k]
            if self.wd != 0:
                g = g + self.wd * p
            self.m[k].mul_(self.b1).add_(g, alpha=1 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            denom = (self.v[k].sqrt() / math.sqrt(bc2)).add_(self.eps)
            p.addcdiv_(self.m[k], denom, value=-self.lr / bc1)


class LazyRowAdam(DenseAdam):
    """Row-sparse variant used by the B200 path for the big tables: identical arithmetic, but a
    table row is updated (moments, weight decay, parameter) only in steps where it appears in the
    batch.  Dense (non-table) parameters follow DenseAdam.  Documented deviation H1."""

    def step(self, params, grads, touched: Optional[Dict[str, torch.Tensor]] = None):
        self.t += 1
        bc1 = 1.0 - self.b1 ** self.t
        bc2 = 1.0 - self.b2 ** self.t
        done = set()
        for k, p in params.items():
            if id(p) in done:
                continue
            done.add(id(p))
            g = grads[k]
            rows = None if touched is None else touched.get(k)
            if rows is None:
                sl = slice(None)
            else:
                rows = torch.unique(rows.long().reshape(-1))
                sl = rows[rows > 0]
            gs, ps = g[sl], p[sl]
            if self.wd != 0:
                gs = gs + self.wd * ps
            m = self.m[k][sl] * self.b1 + (1 - self.b1) * gs
            v = self.v[k][sl] * self.b2 + (1 - self.b2) * gs * gs
            self.m[k][sl] = m
            self.v[k][sl] = v
            p[sl] = ps - (self.lr / bc1) * m / (v.sqrt() / math.sqrt(bc2) + self.eps)


def train_step(model, p, cfg, batch, opt: DenseAdam, drop=None):
    """One iteration of Trainer.fit's inner loop: forward, backward, optimizer step.
    unirec/facility/trainer.py:340-349.  Mutates `p` in place; returns the loss."""
    loss, _, _, grads = loss_and_grads(model, p, cfg, batch, drop=drop)
    if torch.isnan(loss):          # trainer.py:344-352: a NaN loss skips the update
        return loss
    opt.step(p, grads)
    return loss


def clip_grad_norm(grads: Dict[str, torch.Tensor], max_norm: float):
    """torch.nn.utils.clip_grad_norm_ (L2) as used at trainer.py:347-348."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads.values():
        g.mul_(coef)
    return total
