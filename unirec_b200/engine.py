"""Host-side execution engine of the hot path: explicit forward/backward over the CUDA kernels.

One engine per model instance.  It owns
  * the flat fp32 buffer holding every dense (non-table) parameter and the matching flat gradient buffer
    (nn.Parameters are re-pointed into it, state_dict names unchanged -> one optimizer launch, one all-reduce);
  * the activation workspace (allocated once per batch shape, reused every step);
  * per-table row-list state (head/next/uniq) that carries the sparse embedding gradient to the optimizer without
    ever materialising a [V,d] gradient.

Reference call stack being replaced: BaseRecommender.forward -> forward_item_emb / forward_user_emb / _predict_layer /
_cal_loss (unirec/model/base/recommender.py:46-96, reco_abc.py:220-272) and its autograd.
"""
from typing import Dict, List, Optional

import contextlib
import weakref

import torch

from . import ops

# 'fp32' = exact FMA (SIMT); 'tf32x3' = 3xTF32 split on tcgen05 (fp32-class accuracy); 'tf32' = single-pass TF32 tensor cores
PRECISION_CODES = {'fp32': 0, 'tf32': 1, 'bf16': 2, 'tf32x3': 3}

_LO_PLANE_OWNER = None      # id() of the engine whose weight lo plane is registered with the GEMM launcher (Engine.refresh_lo_plane)


def _release_lo_plane_of(owner_id):
    global _LO_PLANE_OWNER
    if _LO_PLANE_OWNER == owner_id:
        _LO_PLANE_OWNER = None
        try:
            ops.gemm_set_lo_plane(None, None)
        except Exception:           # interpreter shutdown: the library may already be gone
            pass


class Workspace:
    """Named device buffers, one per (name, shape, dtype).  A buffer is never freed or moved once handed out: captured CUDA
    graphs (Trainer.train_step) keep raw pointers into it, and train / eval / last-partial-batch shapes alternate."""

    def __init__(self, device):
        self.device = device
        self.buf: Dict[tuple, torch.Tensor] = {}

    def get(self, name, shape, dtype=torch.float32, zero=False):
        shape = tuple(int(s) for s in shape)
        key = (name, shape, dtype)
        t = self.buf.get(key)
        if t is None:
            # zero-filled once: rows beyond the live token count are read (never used) by whole-block kernels and must be finite
            t = torch.zeros(shape, dtype=dtype, device=self.device)
            self.buf[key] = t
        if zero:
            t.zero_()
        return t


class FlatParams:
    """Packs dense parameters into one contiguous fp32 buffer (each slot 16-byte aligned) and re-points
    `param.data` into it.  `groups` lists parameter names that must be adjacent (e.g. query/key/value weights
    -> one [3d,d] matrix)."""

    def __init__(self, named_params: List, device):
        self.names = [n for n, _ in named_params]
        self.offsets, total = {}, 0
        for n, p in named_params:
            self.offsets[n] = (total, p.numel(), tuple(p.shape))
            total += (p.numel() + 3) // 4 * 4
        self.size = total
        self.data = torch.zeros(total, dtype=torch.float32, device=device)
        self.grad = torch.zeros(total, dtype=torch.float32, device=device)
        for n, p in named_params:
            off, cnt, shape = self.offsets[n]
            view = self.data[off:off + cnt].view(shape)
            view.copy_(p.data.to(device))
            p.data = view

    def p(self, name):
        off, cnt, shape = self.offsets[name]
        return self.data[off:off + cnt].view(shape)

    def g(self, name):
        off, cnt, shape = self.offsets[name]
        return self.grad[off:off + cnt].view(shape)

    def span(self, first, last):
        """Contiguous view covering parameters first..last (adjacent in the buffer), as a flat tensor."""
        o0 = self.offsets[first][0]
        o1, c1, _ = self.offsets[last]
        return self.data[o0:o1 + c1], self.grad[o0:o1 + c1]


class RowGrad:
    """Sparse gradient of one table for one step: which batch entries touch which rows, and where the per-entry
    gradient rows live.  Consumed by FusedOptimizer through ur_rowlist_link / ur_rowlist_apply."""

    def __init__(self, param: torch.nn.Parameter):
        self.param = param
        self.head = None           # int32 [V], -1 between steps
        self.next = None
        self.uniq = None
        self.n_uniq = None
        self.n_hist = None         # early-link mode: number of unique rows touched by the history keys (linked first)
        self.early = False         # this step's lists were linked ahead of the backward pass (Engine._early_link)
        self.specs = []            # list of (keys, src, src_group, coef, coef_group)
        self.linked = False
        self.pad_id = 0            # key value that is skipped (the padding id)
        self.shard = (1, 0)        # (world, rank): keys are global ids of a row-sharded table, only owned entries are linked

    def reset(self):
        self.specs = []
        self.linked = False
        self.early = False

    def prepare(self, n):
        """Allocate / size the list state for a step with n entries."""
        dev = self.param.device
        V = self.param.shape[0]
        if self.head is None or self.head.numel() != V:
            self.head = torch.full((V,), -1, dtype=torch.int32, device=dev)
        if self.next is None or self.next.numel() < n:
            if self.next is not None:          # captured CUDA graphs may still point at the smaller buffers: keep them alive
                self.__dict__.setdefault('_retired', []).append((self.next, self.uniq))
            self.next = torch.empty(n, dtype=torch.int32, device=dev)
            self.uniq = torch.empty(n, dtype=torch.int32, device=dev)
        if self.n_uniq is None:
            self.n_uniq = torch.zeros(1, dtype=torch.int32, device=dev)
            self.n_hist = torch.zeros(1, dtype=torch.int32, device=dev)

    def add(self, keys, src, src_group=1, coef=None, coef_group=1, parts=None, key_mask=-1):
        """parts = (int64 device tensor of peer pointers, rows per part): the source rows live in per-rank buffers (sharding.py);
        key_mask strips the flag bits of packed ids (ops.pack_ids)."""
        if len(self.specs) >= 2:
            raise RuntimeError('a table takes at most two gradient sources per step')
        if parts is not None and not self.specs:
            raise RuntimeError('a multi-part source must be the second gradient source of a table')
        self.specs.append((keys, src, int(src_group), coef, int(coef_group), parts, int(key_mask)))

    def n_entries(self):
        return sum(k.numel() for k, *_ in self.specs)

    def link(self):
        """Thread the batch entries onto per-row lists (idempotent within a step)."""
        if self.linked or not self.specs:
            return
        self.prepare(self.n_entries())
        self.n_uniq.zero_()
        off = 0
        for keys, *rest in self.specs:
            ops.rowlist_link(self.head, keys, off, self.next, self.uniq, self.n_uniq, pad_id=self.pad_id, world=self.shard[0],
                             rank=self.shard[1], key_mask=rest[-1])
            off += keys.numel()
        self.linked = True

    def sources(self):
        return [(src, sg, coef, cg, keys.numel(), parts) for keys, src, sg, coef, cg, parts, _mask in self.specs]

    def to_dense(self):
        """Exact-dense mode: materialise the [V,d] gradient the reference's autograd would produce."""
        g = torch.zeros_like(self.param.data)
        if self.shard[0] > 1:
            raise RuntimeError('dense table gradients (table_update: dense) are not available with row-sharded tables')
        for keys, src, sg, coef, cg, parts, _mask in self.specs:
            if parts is not None:
                raise RuntimeError('dense table gradients are not available with peer-memory gradient sources')
            ops.scatter_add_rows(g, keys, src, sg, coef, cg, pad_id=self.pad_id)
        return g


def _lin_fwd(x, M, K, w, b, N, out, act=None, preact=None, lda=None, prec=0):
    """out[M,N] = act(x[M,K] @ w[N,K]^T + b)"""
    return ops.gemm(x, w, out, M, N, K, transB=True, lda=lda, bias=b, act=act, preact=preact, precision=prec)


def _lin_bwd(dy, M, N, x, K, w, dx, dw, db, lda_x=None, accumulate_dx=False, prec=0, lddy=None, lddx=None, dact=None,
             act=None, dx_colsum=None, rows_dev=None, side=None):
    """y = x w^T + b.  dx[M,K] (+)= dy[M,N] @ w[N,K];  dw[N,K] += dy^T x;  db[N] += colsum(dy) (skipped when db is None:
    the producer of dy already accumulated it).  dact/act: dx is multiplied by act'(dact) in the GEMM epilogue, and
    dx_colsum receives the column sums of the result (bias gradient of the layer below).  `side`: a stream for the weight / bias
    gradient, which then runs concurrently with the input-gradient GEMM (B-row products of the trimmed last layer: each launch
    fills a fraction of the GPU); the caller's stream waits for it before this function returns."""
    main = torch.cuda.current_stream() if side is not None else None
    if side is not None:
        side.wait_stream(main)
    if dx is not None:
        if dact is not None or dx_colsum is not None:
            ops.gemm_fused(dy, w, dx, M, K, N, lda=lddy, ldc=lddx, accumulate=accumulate_dx, precision=prec, dact=dact, act=act,
                           colsum=dx_colsum, rows_dev=rows_dev)
        else:
            ops.gemm(dy, w, dx, M, K, N, lda=lddy, ldc=lddx, accumulate=accumulate_dx, precision=prec, rows_dev=rows_dev)
    with torch.cuda.stream(side) if side is not None else contextlib.nullcontext():
        ops.gemm(dy, x, dw, N, K, M, transA=True, lda=lddy or N, ldb=lda_x, accumulate=True, precision=prec, rows_dev=rows_dev)
        if db is not None:
            ops.colsum_accum(dy, M, N, db, ldx=lddy, rows_dev=rows_dev)
    if side is not None:
        main.wait_stream(side)


# ==================================================================================================
# Towers: forward(batch) -> user_emb [B,d];  backward(d_user_emb) -> dense grads into FlatParams.grad, row grads
# registered on the tables' RowGrad.
# ==================================================================================================
class SASRecTower:
    """unirec/model/sequential/sasrec.py:59-76 over modules.TransformerEncoder (modules.py:247-433).

    Token layout: the encoder runs on the LIVE positions only (csrc/pack.cu: real items + position L-1, every position of a
    sequence without real items), packed sample by sample into the first n_tok rows of [B*L, .] buffers.  n_tok stays on the
    device: launches are sized for B*L rows and every kernel stops at n_tok (`rows_dev`), so the step needs no host
    synchronisation and remains CUDA-graph capturable.  `pack_sequences: 0` uses the identity map (all B*L positions)."""

    def __init__(self, eng, cfg):
        self.eng = eng
        self.d = int(cfg['embedding_size'])
        self.n_layers = int(cfg['n_layers'])
        self.H = int(cfg['n_heads'])
        self.I = int(cfg['inner_size'])
        self.act = cfg['hidden_act']
        self.eps = float(cfg['layer_norm_eps'])
        self.causal = bool(cfg['use_position_emb'])
        self.dh = self.d // self.H
        self.trim_last = bool(int(cfg.get('trim_last_layer', 1)))    # 0: compute the dead rows of the last layer too (A/B testing)
        self.packed = bool(int(cfg.get('pack_sequences', 1)))        # 0: keep every position (identity token map)
        # nn.Dropout sites of the reference (sasrec.py:69; modules.py:307,313,352), regenerated by counter in the backward kernels
        self.p_hidden = float(cfg.get('hidden_dropout_prob', 0) or 0)
        self.p_attn = float(cfg.get('attn_dropout_prob', 0) or 0)
        for p_ in (self.p_hidden, self.p_attn):
            if not 0.0 <= p_ < 1.0:
                raise ValueError('dropout probabilities must lie in [0, 1), got %r' % (p_,))

    # dropout site ids: 0 = input; layer i: 1+3i attention probabilities, 2+3i attention output, 3+3i FFN output
    def _drop(self, site, p, row_pos=None, pos_mul=1, pos_add=0):
        if p <= 0.0 or not self.eng.model.training:
            return None
        return ops.Drop(self.eng.rng, p, site, row_pos, pos_mul, pos_add)

    @staticmethod
    def flat_order(model):
        names = []
        if model.position_embedding is not None:
            names.append('position_embedding.weight')
        names += ['LayerNorm.weight', 'LayerNorm.bias']
        for i in range(model.n_layers):
            a, f = 'trm_encoder.layer.%d.multi_head_attention.' % i, 'trm_encoder.layer.%d.feed_forward.' % i
            names += [a + 'query.weight', a + 'key.weight', a + 'value.weight',
                      a + 'query.bias', a + 'key.bias', a + 'value.bias',
                      a + 'dense.weight', a + 'dense.bias', a + 'LayerNorm.weight', a + 'LayerNorm.bias',
                      f + 'dense_1.weight', f + 'dense_1.bias', f + 'dense_2.weight', f + 'dense_2.bias',
                      f + 'LayerNorm.weight', f + 'LayerNorm.bias']
        return names

    def _names(self, i):
        return 'trm_encoder.layer.%d.multi_head_attention.' % i, 'trm_encoder.layer.%d.feed_forward.' % i

    def _pack(self, item_seq):
        ws = self.eng.ws
        B, L = item_seq.shape
        pk = dict(offs=ws.get('pk_offs', (B + 1,), torch.int32), tok_src=ws.get('pk_src', (B * L,), torch.int32),
                  tok_inv=ws.get('pk_inv', (B * L,), torch.int32), last=ws.get('pk_last', (B,), torch.int32),
                  n=ws.get('pk_n', (1,), torch.int32))
        ops.pack_tokens(item_seq, pk['offs'], pk['tok_src'], pk['tok_inv'], pk['last'], pk['n'], keep_all=not self.packed)
        return pk

    # ---------------------------------------------------------------- forward
    def forward(self, item_seq, save=True, **_):
        eng, fp, ws = self.eng, self.eng.flat, self.eng.ws
        eng.refresh_lo_plane()
        B, L = item_seq.shape
        d, T = self.d, B * L
        item_seq = item_seq.contiguous()
        pk = self._pack(item_seq)
        table, index = eng.seq_rows_source(item_seq)
        self.seq_src = (table, index)
        pos = fp.p('position_embedding.weight') if self.causal else None
        x = ws.get('x0', (T, d))
        self.mean0, self.rstd0 = ws.get('mean0', (T,)), ws.get('rstd0', (T,))
        self.drop0 = self._drop(0, self.p_hidden)
        ops.seq_prep_ln_fwd(table, pos, fp.p('LayerNorm.weight'), fp.p('LayerNorm.bias'), self.eps, index, x,
                            self.mean0, self.rstd0, tok_src=pk['tok_src'], n_tok=pk['n'], shards=eng.seq_shards, drop=self.drop0)
        self.saved = []
        self.item_seq, self.pk = item_seq, pk
        user = ws.get('user_emb', (B, d))
        for i in range(self.n_layers):
            tag = str(i) if save else 'e'
            if i == self.n_layers - 1 and self.trim_last:
                st = self._layer_fwd_last(i, x, item_seq, pk, tag, user)
            else:
                st = self._layer_fwd_full(i, x, item_seq, pk, tag)
            if save:
                self.saved.append(st)
            x = st[-2]                  # layer output (the tuple ends with the layer's dropout sites)
        if not self.trim_last:
            ops.gather_rows(x, pk['last'], out=user)
        return user

    def _layer_fwd_full(self, i, x, item_seq, pk, tag):
        eng, fp, ws = self.eng, self.eng.flat, self.eng.ws
        B, L = item_seq.shape
        d, I, H, T, prec, n = self.d, self.I, self.H, B * L, eng.prec, pk['n']
        a, f = self._names(i)
        wqkv, _ = fp.span(a + 'query.weight', a + 'value.weight')
        bqkv, _ = fp.span(a + 'query.bias', a + 'value.bias')
        qkv = ws.get('qkv' + tag, (T, 3 * d))
        ops.gemm(x, wqkv, qkv, T, 3 * d, d, transB=True, bias=bqkv, precision=prec, rows_dev=n)
        ctx, lse = ws.get('ctx' + tag, (T, d)), ws.get('lse' + tag, (B, H, L))
        drops = (self._drop(1 + 3 * i, self.p_attn), self._drop(2 + 3 * i, self.p_hidden, pk['tok_src']),
                 self._drop(3 + 3 * i, self.p_hidden, pk['tok_src']))
        ops.attn_fwd(qkv, item_seq, H, self.dh, self.causal, ctx, lse, offs=pk['offs'], tok_src=pk['tok_src'], drop=drops[0])
        z1 = ws.get('z1' + tag, (T, d))
        ops.gemm(ctx, fp.p(a + 'dense.weight'), z1, T, d, d, transB=True, bias=fp.p(a + 'dense.bias'), precision=prec, rows_dev=n)
        x1 = ws.get('x1' + tag, (T, d))
        m1, r1 = ws.get('m1' + tag, (T,)), ws.get('r1' + tag, (T,))
        ops.add_ln_fwd(z1, x, fp.p(a + 'LayerNorm.weight'), fp.p(a + 'LayerNorm.bias'), self.eps, x1, m1, r1, rows_dev=n,
                       drop=drops[1])
        hpre, hact = ws.get('hpre' + tag, (T, I)), ws.get('hact' + tag, (T, I))
        ops.gemm(x1, fp.p(f + 'dense_1.weight'), hact, T, I, d, transB=True, bias=fp.p(f + 'dense_1.bias'), act=self.act,
                 preact=hpre, precision=prec, rows_dev=n)
        z2 = ws.get('z2' + tag, (T, d))
        ops.gemm(hact, fp.p(f + 'dense_2.weight'), z2, T, d, I, transB=True, bias=fp.p(f + 'dense_2.bias'), precision=prec,
                 rows_dev=n)
        x2 = ws.get('x2' + tag, (T, d))
        m2, r2 = ws.get('m2' + tag, (T,)), ws.get('r2' + tag, (T,))
        ops.add_ln_fwd(z2, x1, fp.p(f + 'LayerNorm.weight'), fp.p(f + 'LayerNorm.bias'), self.eps, x2, m2, r2, rows_dev=n,
                       drop=drops[2])
        return ('full', x, qkv, ctx, lse, z1, m1, r1, x1, hpre, hact, z2, m2, r2, x2, drops)

    def _layer_fwd_last(self, i, x, item_seq, pk, tag, user):
        """Last encoder layer: only position L-1 of its output reaches the scorer (sasrec.py:74-75), so everything after the
        key/value projection is computed for the B last rows only (compact [B, .] buffers).  Keys and values still cover every
        live position.  Results are identical to the full computation: the other output rows are dead values in the reference."""
        eng, fp, ws = self.eng, self.eng.flat, self.eng.ws
        B, L = item_seq.shape
        d, I, H, T, prec, n = self.d, self.I, self.H, B * L, eng.prec, pk['n']
        a, f = self._names(i)
        wqkv, _ = fp.span(a + 'query.weight', a + 'value.weight')
        bqkv, _ = fp.span(a + 'query.bias', a + 'value.bias')
        qkv = ws.get('qkv' + tag, (T, 3 * d))
        # K | V for every live position (rows d..3d of the packed weight), Q for the last position only
        ops.gemm(x, wqkv[d * d:], qkv[:, d:], T, 2 * d, d, transB=True, ldc=3 * d, bias=bqkv[d:], precision=prec, rows_dev=n)
        xl = ops.gather_rows(x, pk['last'], out=ws.get('xl' + tag, (B, d)))
        ql = ws.get('ql' + tag, (B, d))
        ops.gemm(xl, wqkv[:d * d], ql, B, d, d, transB=True, bias=bqkv[:d], precision=prec)
        ctx, lse = ws.get('ctxl' + tag, (B, d)), ws.get('lse' + tag, (B, H, L))
        # compact rows: row b is position (b, L-1)
        drops = (self._drop(1 + 3 * i, self.p_attn), self._drop(2 + 3 * i, self.p_hidden, None, L, L - 1),
                 self._drop(3 + 3 * i, self.p_hidden, None, L, L - 1))
        ops.attn_fwd(qkv, item_seq, H, self.dh, self.causal, ctx, lse, q_only_last=True, offs=pk['offs'],
                     tok_src=pk['tok_src'], q_last=ql, drop=drops[0])
        z1 = ws.get('z1l' + tag, (B, d))
        ops.gemm(ctx, fp.p(a + 'dense.weight'), z1, B, d, d, transB=True, bias=fp.p(a + 'dense.bias'), precision=prec)
        x1 = ws.get('x1l' + tag, (B, d))
        m1, r1 = ws.get('m1l' + tag, (B,)), ws.get('r1l' + tag, (B,))
        ops.add_ln_fwd(z1, xl, fp.p(a + 'LayerNorm.weight'), fp.p(a + 'LayerNorm.bias'), self.eps, x1, m1, r1, rows=B, d=d,
                       drop=drops[1])
        hpre, hact = ws.get('hprel' + tag, (B, I)), ws.get('hactl' + tag, (B, I))
        _lin_fwd(x1, B, d, fp.p(f + 'dense_1.weight'), fp.p(f + 'dense_1.bias'), I, hact, act=self.act, preact=hpre, prec=prec)
        z2 = ws.get('z2l' + tag, (B, d))
        _lin_fwd(hact, B, I, fp.p(f + 'dense_2.weight'), fp.p(f + 'dense_2.bias'), d, z2, prec=prec)
        m2, r2 = ws.get('m2l' + tag, (B,)), ws.get('r2l' + tag, (B,))
        ops.add_ln_fwd(z2, x1, fp.p(f + 'LayerNorm.weight'), fp.p(f + 'LayerNorm.bias'), self.eps, user, m2, r2, rows=B, d=d,
                       drop=drops[2])
        return ('last', x, qkv, ctx, lse, z1, m1, r1, x1, hpre, hact, z2, m2, r2, xl, ql, user, drops)

    def _wgrad_stream(self):
        """Side stream for the weight-gradient products of the trimmed last layer (B rows each: a fraction of the GPU per launch)."""
        if getattr(self, '_wside', None) is None:
            self._wside = torch.cuda.Stream(device=self.eng.device)
        return self._wside

    # ---------------------------------------------------------------- backward
    def backward(self, d_user):
        eng, fp, ws = self.eng, self.eng.flat, self.eng.ws
        item_seq, pk = self.item_seq, self.pk
        B, L = item_seq.shape
        d, T = self.d, B * L
        dx = None
        for i in reversed(range(self.n_layers)):
            st = self.saved[i]
            if st[0] == 'last':
                dx = self._layer_bwd_last(i, st, d_user, item_seq, pk)
            else:
                if dx is None:                      # untrimmed last layer: the loss gradient enters at position L-1
                    dx = ws.get('dx_a', (T, d), zero=True)
                    ops.scatter_add_rows(dx, pk['last'], d_user, pad_id=-1)
                dx = self._layer_bwd_full(i, st, dx, item_seq, pk)
        table, index = self.seq_src
        pos = fp.p('position_embedding.weight') if self.causal else None
        drows = ws.get('drows', (T, d))
        ops.seq_prep_ln_bwd(table, pos, fp.p('LayerNorm.weight'), index, self.mean0, self.rstd0, dx, drows,
                            fp.g('LayerNorm.weight'), fp.g('LayerNorm.bias'),
                            fp.g('position_embedding.weight') if self.causal else None, tok_inv=pk['tok_inv'],
                            shards=eng.seq_shards, drop=self.drop0)
        eng.add_seq_rowgrad(item_seq, drows)

    def _layer_bwd_full(self, i, st, dx, item_seq, pk):
        eng, fp, ws = self.eng, self.eng.flat, self.eng.ws
        B, L = item_seq.shape
        d, I, H, T, prec, n = self.d, self.I, self.H, B * L, eng.prec, pk['n']
        a, f = self._names(i)
        _, x, qkv, ctx, lse, z1, m1, r1, x1, hpre, hact, z2, m2, r2, _x2, drops = st
        # x2 = LN(z2), z2 = dropout(ffn(x1)) + x1.  Bias gradients ride on the kernels that produce the corresponding dy.  Every
        # token-reduction GEMM (dW = dy^T x) has one operand written by an LN kernel, whose rows [n_tok, roundup32) are zero.
        # With dropout the LN backward emits two gradients: dz (residual branch) and dz * mask (linear branch, `dy`).
        dz2 = ws.get('dz2', (T, d))
        dy2 = ws.get('dz2m', (T, d)) if drops[2] is not None else dz2
        ops.add_ln_bwd(z2, fp.p(f + 'LayerNorm.weight'), m2, r2, dx, dz2, fp.g(f + 'LayerNorm.weight'),
                       fp.g(f + 'LayerNorm.bias'), dzsum=fp.g(f + 'dense_2.bias'), rows_dev=n, drop=drops[2], dZdrop=dy2)
        dh = ws.get('dh', (T, I))
        # dh = (dy2 @ W2) * act'(hpre), db1 += colsum(dh): one GEMM with a fused epilogue
        _lin_bwd(dy2, T, d, hact, I, fp.p(f + 'dense_2.weight'), dh, fp.g(f + 'dense_2.weight'), None, prec=prec,
                 dact=hpre, act=self.act, dx_colsum=fp.g(f + 'dense_1.bias'), rows_dev=n)
        # dx1 = dz2 (residual) + dh @ W1   -> accumulate into dz2
        _lin_bwd(dh, T, I, x1, d, fp.p(f + 'dense_1.weight'), dz2, fp.g(f + 'dense_1.weight'), None,
                 accumulate_dx=True, prec=prec, rows_dev=n)
        # x1 = LN(z1), z1 = attn_out + x
        dz1 = ws.get('dz1_%d' % (i % 2), (T, d))     # becomes this layer's input gradient (no copy)
        dy1 = ws.get('dz1m', (T, d)) if drops[1] is not None else dz1
        ops.add_ln_bwd(z1, fp.p(a + 'LayerNorm.weight'), m1, r1, dz2, dz1, fp.g(a + 'LayerNorm.weight'),
                       fp.g(a + 'LayerNorm.bias'), dzsum=fp.g(a + 'dense.bias'), rows_dev=n, drop=drops[1], dZdrop=dy1)
        dctx = ws.get('dctx', (T, d))
        _lin_bwd(dy1, T, d, ctx, d, fp.p(a + 'dense.weight'), dctx, fp.g(a + 'dense.weight'), None, prec=prec, rows_dev=n)
        dqkv = ws.get('dqkv', (T, 3 * d))
        ops.attn_bwd(qkv, item_seq, H, self.dh, self.causal, ctx, lse, dctx, dqkv, offs=pk['offs'], tok_src=pk['tok_src'],
                     drop=drops[0])
        wqkv, gwqkv = fp.span(a + 'query.weight', a + 'value.weight')
        _, gbqkv = fp.span(a + 'query.bias', a + 'value.bias')
        # dx = dz1 (residual) + dqkv @ Wqkv  -> accumulate into dz1
        _lin_bwd(dqkv, T, 3 * d, x, d, wqkv, dz1, gwqkv, gbqkv, accumulate_dx=True, prec=prec, rows_dev=n)
        return dz1

    def _layer_bwd_last(self, i, st, d_user, item_seq, pk):
        """Backward of _layer_fwd_last: the loss gradient exists at position L-1 only, so everything down to the attention is
        B rows; dK/dV (all live positions) and dQ (last position) then give the full-size input gradient."""
        eng, fp, ws = self.eng, self.eng.flat, self.eng.ws
        B, L = item_seq.shape
        d, I, H, T, prec, n = self.d, self.I, self.H, B * L, eng.prec, pk['n']
        a, f = self._names(i)
        _, x, qkv, ctx, lse, z1, m1, r1, x1, hpre, hact, z2, m2, r2, xl, ql, _user, drops = st
        dz2 = ws.get('dz2l', (B, d))
        dy2 = ws.get('dz2ml', (B, d)) if drops[2] is not None else dz2
        ops.add_ln_bwd(z2, fp.p(f + 'LayerNorm.weight'), m2, r2, d_user, dz2, fp.g(f + 'LayerNorm.weight'),
                       fp.g(f + 'LayerNorm.bias'), rows=B, d=d, dzsum=fp.g(f + 'dense_2.bias'), drop=drops[2], dZdrop=dy2)
        dh = ws.get('dhl', (B, I))
        side = self._wgrad_stream()
        _lin_bwd(dy2, B, d, hact, I, fp.p(f + 'dense_2.weight'), dh, fp.g(f + 'dense_2.weight'), None, prec=prec,
                 dact=hpre, act=self.act, dx_colsum=fp.g(f + 'dense_1.bias'), side=side)
        _lin_bwd(dh, B, I, x1, d, fp.p(f + 'dense_1.weight'), dz2, fp.g(f + 'dense_1.weight'), None,
                 accumulate_dx=True, prec=prec, side=side)
        dz1 = ws.get('dz1l', (B, d))
        dy1 = ws.get('dz1ml', (B, d)) if drops[1] is not None else dz1
        ops.add_ln_bwd(z1, fp.p(a + 'LayerNorm.weight'), m1, r1, dz2, dz1, fp.g(a + 'LayerNorm.weight'),
                       fp.g(a + 'LayerNorm.bias'), rows=B, d=d, dzsum=fp.g(a + 'dense.bias'), drop=drops[1], dZdrop=dy1)
        dctx = ws.get('dctxl', (B, d))
        _lin_bwd(dy1, B, d, ctx, d, fp.p(a + 'dense.weight'), dctx, fp.g(a + 'dense.weight'), None, prec=prec, side=side)
        dqkv, dql = ws.get('dqkv', (T, 3 * d)), ws.get('dql', (B, d))
        ops.attn_bwd(qkv, item_seq, H, self.dh, self.causal, ctx, lse, dctx, dqkv, q_only_last=True, offs=pk['offs'],
                     tok_src=pk['tok_src'], q_last=ql, dq_last=dql, drop=drops[0])
        dkv = dqkv[:, d:]                                                  # [T, 2d], row stride 3d
        wqkv, gwqkv = fp.span(a + 'query.weight', a + 'value.weight')
        _, gbqkv = fp.span(a + 'query.bias', a + 'value.bias')
        wq, wkv, gwq, gwkv = wqkv[:d * d], wqkv[d * d:], gwqkv[:d * d], gwqkv[d * d:]
        # input gradient: every live row gets dKV @ Wkv; the last rows add dQ @ Wq and the residual branch (dz1)
        dxf = ws.get('dz1_%d' % (i % 2), (T, d))
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            # weight / bias gradients of the projections (x comes from an LN kernel: zero rows up to the next k-block), next to
            # the input-gradient products below
            ops.gemm(dql, xl, gwq, d, d, B, transA=True, lda=d, ldb=d, accumulate=True, precision=prec)
            ops.colsum_accum(dql, B, d, gbqkv[:d])
            ops.colsum_accum(dkv, T, 2 * d, gbqkv[d:], ldx=3 * d, rows_dev=n)
        ops.gemm(dkv, wkv, dxf, T, d, 2 * d, lda=3 * d, precision=prec, rows_dev=n)
        ops.gemm(dql, wq, dz1, B, d, d, accumulate=True, precision=prec)        # dz1 <- dz1 + dQ @ Wq
        ops.scatter_add_rows(dxf, pk['last'], dz1, pad_id=-1)
        ops.gemm(dkv, x, gwkv, 2 * d, d, T, transA=True, lda=3 * d, ldb=d, accumulate=True, precision=prec, rows_dev=n)
        main.wait_stream(side)
        return dxf


class GRUTower:
    """unirec/model/sequential/gru.py:27-35.  `dense` is applied to the last step only (the other L-1 outputs of
    the reference's all-steps Linear are dead)."""

    def __init__(self, eng, cfg):
        self.eng = eng
        self.d = int(cfg['embedding_size'])
        self.Hd = int(cfg.get('hidden_size', self.d))
        self.n_groups = int(cfg.get('gru_row_groups', 1))           # per-step path: row groups on parallel streams (measured at c3:
                                                                    # 1 -> 9.28 ms, 4 -> 9.18 ms, 8 -> 12.0 ms per step: no gain, default serial)
        self.persistent = bool(int(cfg.get('gru_persistent', 0)))   # 1: one launch per direction for all L steps (csrc/gru.cu); 0: one
                                                                    # tcgen05 GEMM + gate kernel per time step inside the step's CUDA graph (faster at c3 today)
        self.p_emb = float(cfg.get('dropout_prob', 0) or 0)        # nn.Dropout on the gathered rows, gru.py:29
        if not 0.0 <= self.p_emb < 1.0:
            raise ValueError('dropout_prob must lie in [0, 1), got %r' % (self.p_emb,))

    @staticmethod
    def flat_order(model):
        return ['gru_layers.weight_ih_l0', 'gru_layers.weight_hh_l0', 'gru_layers.bias_ih_l0', 'gru_layers.bias_hh_l0',
                'dense.weight', 'dense.bias']

    # ---- row groups of the per-step recurrence on parallel streams ----
    def _row_groups(self, B):
        """Contiguous row ranges (multiples of 128 rows = one GEMM stripe) whose recurrences run concurrently."""
        G = max(1, min(self.n_groups, B // 256))
        per = ((B + G - 1) // G + 127) // 128 * 128
        groups = [(b0, min(B, b0 + per)) for b0 in range(0, B, per)]
        while len(getattr(self, '_streams', [])) < len(groups):
            self.__dict__.setdefault('_streams', []).append(torch.cuda.Stream(device=self.eng.device))
        return groups

    def _group_streams(self, groups):
        return [(g, self._streams[i]) for i, g in enumerate(groups)]

    def _fork(self, groups):
        main = torch.cuda.current_stream()
        for i in range(len(groups)):
            self._streams[i].wait_stream(main)

    def _join(self, groups):
        main = torch.cuda.current_stream()
        for i in range(len(groups)):
            main.wait_stream(self._streams[i])

    def forward(self, item_seq, save=True, **_):
        eng, fp, ws = self.eng, self.eng.flat, self.eng.ws
        eng.refresh_lo_plane()
        B, L = item_seq.shape
        d, Hd, prec = self.d, self.Hd, eng.prec
        table, index = eng.seq_rows_source(item_seq)
        x = ws.get('gru_x', (B * L, d))
        ops.gather_rows(table, index, out=x)
        self.drop0 = ops.Drop(eng.rng, self.p_emb, 0) if (self.p_emb > 0 and eng.model.training) else None
        ops.dropout_rows(x, self.drop0)
        gi = ws.get('gru_gi', (B * L, 3 * Hd))
        _lin_fwd(x, B * L, d, fp.p('gru_layers.weight_ih_l0'), fp.p('gru_layers.bias_ih_l0'), 3 * Hd, gi, prec=prec)
        hs = ws.get('gru_h', (L + 1, B, Hd))
        hs[0].zero_()
        save_g = ws.get('gru_save', (L, B, 4 * Hd))
        gh = ws.get('gru_gh', (B, 3 * Hd))
        w_hh, b_hh = fp.p('gru_layers.weight_hh_l0'), fp.p('gru_layers.bias_hh_l0')
        if self.persistent and Hd % 32 == 0 and Hd <= 768:
            # one launch for all L steps: W_hh^T streamed from L2, hidden tile resident in shared memory (csrc/gru.cu)
            whh_t = ops.transpose(w_hh, ws.get('gru_whh_t', (Hd, 3 * Hd)))
            ops.gru_seq_fwd(gi, whh_t, b_hh, hs, save_g, B, L, Hd)
        else:
            # Per-step path.  Batch rows are independent along the recurrence and every per-step kernel is latency-bound (a
            # [B, 3h] x h product gives the persistent GEMM ~100 CTAs for ~15 us), so the batch is cut into row groups whose
            # step chains run on parallel streams -- parallel branches of the step's CUDA graph.
            groups = self._row_groups(B)
            gi3 = gi.view(B, L, 3 * Hd)
            self._fork(groups)
            for t in range(L):
                for gidx, (b0, b1) in enumerate(groups):
                    with torch.cuda.stream(self._streams[gidx]):
                        n = b1 - b0
                        _lin_fwd(hs[t][b0:b1], n, Hd, w_hh, b_hh, 3 * Hd, gh[b0:b1], prec=prec)
                        ops.gru_gate_fwd(gi3[b0:b1, t], L * 3 * Hd, gh[b0:b1], hs[t][b0:b1], hs[t + 1][b0:b1], save_g[t][b0:b1], n, Hd)
            self._join(groups)
        user = ws.get('user_emb', (B, d))
        _lin_fwd(hs[L], B, Hd, fp.p('dense.weight'), fp.p('dense.bias'), d, user, prec=prec)
        self.item_seq, self.x, self.hs, self.save_g = item_seq, x, hs, save_g
        return user

    def backward(self, d_user):
        eng, fp, ws = self.eng, self.eng.flat, self.eng.ws
        item_seq = self.item_seq
        B, L = item_seq.shape
        d, Hd, prec = self.d, self.Hd, eng.prec
        hs, save_g = self.hs, self.save_g
        dh = ws.get('gru_dh_a', (B, Hd))
        _lin_bwd(d_user, B, d, hs[L], Hd, fp.p('dense.weight'), dh, fp.g('dense.weight'), fp.g('dense.bias'), prec=prec)
        dgi = ws.get('gru_dgi', (B * L, 3 * Hd))
        dgh_all = ws.get('gru_dgh', (L, B, 3 * Hd))
        w_hh = fp.p('gru_layers.weight_hh_l0')
        if self.persistent and Hd % 32 == 0 and Hd <= 768:
            ops.gru_seq_bwd(dh, save_g, hs, w_hh, dgi, dgh_all, B, L, Hd)
        else:
            groups = self._row_groups(B)
            dgi3 = dgi.view(B, L, 3 * Hd)
            dh_a, dh_b = dh, ws.get('gru_dh_b', (B, Hd))
            self._fork(groups)
            for t in reversed(range(L)):
                dh_prev = dh_b if dh is dh_a else dh_a
                for gidx, (b0, b1) in enumerate(groups):
                    with torch.cuda.stream(self._streams[gidx]):
                        n = b1 - b0
                        ops.gru_gate_bwd(dh[b0:b1], save_g[t][b0:b1], hs[t][b0:b1], dgi3[b0:b1, t], L * 3 * Hd, dgh_all[t][b0:b1],
                                         dh_prev[b0:b1], n, Hd)
                        ops.gemm(dgh_all[t][b0:b1], w_hh, dh_prev[b0:b1], n, Hd, 3 * Hd, accumulate=True, precision=prec)
                dh = dh_prev
            self._join(groups)
        # weight gradients over all steps at once
        ops.gemm(dgh_all, hs, fp.g('gru_layers.weight_hh_l0'), 3 * Hd, Hd, L * B, transA=True, lda=3 * Hd, ldb=Hd,
                 accumulate=True, precision=prec)
        ops.colsum_accum(dgh_all, L * B, 3 * Hd, fp.g('gru_layers.bias_hh_l0'))
        drows = ws.get('drows', (B * L, d))
        _lin_bwd(dgi, B * L, 3 * Hd, self.x, d, fp.p('gru_layers.weight_ih_l0'), drows, fp.g('gru_layers.weight_ih_l0'),
                 fp.g('gru_layers.bias_ih_l0'), prec=prec)
        ops.dropout_rows(drows, self.drop0)
        eng.add_seq_rowgrad(item_seq, drows)


class PoolTower:
    """AvgHist (avghist.py:34-42), SVD++ (svdplusplus.py:31-39) and MF (recommender.py:42-44): sum-pool of the history
    rows (+ user row).  Backward produces no activations at all: the row gradient of history item (b,l) is
    coeff[b] * d_user[b], expressed as a (source row, coefficient) pair for the row-sparse optimizer."""

    def __init__(self, eng, cfg, use_seq, use_user):
        self.eng, self.use_seq, self.use_user = eng, use_seq, use_user
        self.alpha = float(cfg.get('user_sequence_alpha', 0.5))
        self.d = int(cfg['embedding_size'])

    @staticmethod
    def flat_order(model):
        return []

    def forward(self, item_seq=None, item_seq_len=None, user_id=None, save=True, **_):
        return self.eng.pool_forward(item_seq, item_seq_len, user_id, self.use_seq, self.use_user, self.alpha)

    def backward(self, d_user):
        self.eng.pool_backward(d_user, self.use_seq, self.use_user)


# ==================================================================================================
class Engine:
    def __init__(self, model, tower_kind: str):
        self.model = model
        self.cfg = model.config
        self.device = None
        self.flat: Optional[FlatParams] = None
        self.ws: Optional[Workspace] = None
        self.tower_kind = tower_kind
        self.tower = None
        self.prec = PRECISION_CODES[str(self.cfg.get('gemm_precision', 'fp32'))]
        self._rowgrads: Dict[int, RowGrad] = {}
        self.dense_table_grads = False      # exact-dense mode (reference autograd semantics)
        self.nan_flag = None
        self.last = None
        # stream-level overlap of the table update with the encoder backward (set up by the Trainer, see _early_link)
        self.seq_shards = None            # (peer pointers, W) when the history table is row-sharded and read over NVLink (sharding.py)
        self.overlap_hook = None          # FusedOptimizer (provides early_apply) or None
        self.overlap_mode = 1             # 1: link the row lists on a side stream during the forward pass;
                                          # 2: also update the target-only rows on the side stream during the backward pass
        self._side = None
        self._ev_main = self._ev_link = None

    # ---- setup ------------------------------------------------------------------------------
    def table_for_seq(self):
        m = self.model
        return m.item_dst_embedding.weight if hasattr(m, 'item_dst_embedding') else m.item_embedding.weight

    def table_for_target(self):
        return self.model.item_embedding.weight

    def table_params(self):
        m = self.model
        out, seen = [], set()
        for name in ('user_embedding', 'item_embedding', 'item_dst_embedding'):
            mod = getattr(m, name, None)
            if mod is not None and id(mod.weight) not in seen:
                seen.add(id(mod.weight))
                out.append(mod.weight)
        return out

    def link_filter(self):
        """(world, rank) of the row partition of the tables (keys of the row lists are global ids)."""
        return 1, 0

    def rowgrad(self, param) -> RowGrad:
        rg = self._rowgrads.get(id(param))
        if rg is None:
            rg = self._rowgrads[id(param)] = RowGrad(param)
            rg.shard = self.link_filter()
        return rg

    def rowgrads(self):
        return [self.rowgrad(p) for p in self.table_params()]

    def ensure_ready(self):
        dev = self.model.item_embedding.weight.device
        if dev.type != 'cuda':
            raise RuntimeError('unirec_b200 models run on CUDA (sm_100a) only: the hot path is hand-written CUDA with no '
                               'CPU fallback. Move the model with .to("cuda") / config device.')
        if self.device == dev and self.flat is not None:
            return
        self.device = dev
        m, cfg = self.model, self.cfg
        if self.tower_kind == 'sasrec':
            self.tower = SASRecTower(self, cfg)
            order = SASRecTower.flat_order(m)
        elif self.tower_kind == 'gru':
            self.tower = GRUTower(self, cfg)
            order = GRUTower.flat_order(m)
        else:
            self.tower = PoolTower(self, cfg, use_seq=self.tower_kind in ('avghist', 'svdpp'),
                                   use_user=self.tower_kind in ('svdpp', 'mf'))
            order = []
        named = dict(m.named_parameters())
        table_ids = {id(p) for p in self.table_params()}
        rest = [n for n, p in named.items() if n not in order and id(p) not in table_ids]
        self.dense_names = order + rest        # user_bias / item_bias (and anything else) follow the tower params
        self.flat = FlatParams([(n, named[n]) for n in self.dense_names], dev)
        # the tower weights lead the buffer: [0, _lo_span) is what the 3xTF32 lo plane covers (refresh_lo_plane)
        self._lo_span = max([self.flat.offsets[n][0] + (self.flat.offsets[n][1] + 3) // 4 * 4 for n in order], default=0)
        self.flat_lo = None
        self._lo_finalizer = getattr(self, '_lo_finalizer', None)
        for p in self.table_params():
            if not p.data.is_contiguous():
                p.data = p.data.contiguous()
        self.ws = Workspace(dev)
        self.nan_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self._rowgrads = {}
        # dropout counter state (csrc/dropout.cuh): (seed, training step).  The step advances on the device at the start of every
        # training forward_loss, so a replayed CUDA graph draws fresh masks; ranks draw from different seeds.
        seed = int(cfg.get('seed', 2022) or 0) + 0x9E3779B97F4A7C15 * int(getattr(m, 'shard_rank', 0) or 0)
        self.rng = torch.tensor([seed & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64, device=dev)
        self.uses_dropout = any(float(cfg.get(k, 0) or 0) > 0 for k in
                                (('hidden_dropout_prob', 'attn_dropout_prob') if self.tower_kind == 'sasrec' else
                                 ('dropout_prob',) if self.tower_kind == 'gru' else ()))

    def refresh_lo_plane(self):
        """gemm_precision tf32x3: recompute the lo plane of the tower weights (lo = tf32(w - trunc_tf32(w))) and register it with
        the GEMM launcher, which then fetches the lo term of every weight operand by TMA instead of splitting the staged weight tile
        in each CTA of each GEMM (csrc/gemm_tc.cu).  Called at the start of every tower forward: weights only change between steps
        (optimizer, load_state_dict, broadcast), so forward and backward of a step see a consistent plane."""
        if self.prec != PRECISION_CODES['tf32x3'] or not self._lo_span:
            return
        if self.flat_lo is None:
            self.flat_lo = torch.empty(self._lo_span, dtype=torch.float32, device=self.device)
        w = self.flat.data[:self._lo_span]
        ops.split_lo(w, self.flat_lo)
        ops.gemm_set_lo_plane(w, self.flat_lo)
        # The registration is a process-wide (address range -> plane) pair: it is scoped to this forward / backward pair
        # (release_lo_plane) and dropped with the engine, so a later tensor that happens to be allocated at these addresses
        # can never be matched against this engine's plane.
        global _LO_PLANE_OWNER
        if self._lo_finalizer is None:
            self._lo_finalizer = weakref.finalize(self, _release_lo_plane_of, id(self))
        _LO_PLANE_OWNER = id(self)

    def release_lo_plane(self):
        """End of the window opened by refresh_lo_plane (after the tower backward, or after an inference forward)."""
        _release_lo_plane_of(id(self))

    def set_dropout_state(self, seed, next_step):
        """The next training forward draws the masks of (seed, next_step) -- tests replay a known mask set."""
        self.ensure_ready()
        self.rng.copy_(torch.tensor([int(seed), int(next_step) - 1], dtype=torch.int64))

    # ---- overlap of the row-sparse table update with the encoder (single-table sequence towers) ----
    def _overlap_rowgrad(self):
        """The RowGrad eligible for early linking, or None.  Eligible: a Trainer enabled it, plain (unsharded) engine, SASRec/GRU
        tower (history and targets index the same table), row-sparse updates."""
        if self.overlap_hook is None or type(self) is not Engine or self.tower_kind not in ('sasrec', 'gru'):
            return None
        if self.model.table_update == 'dense' or not self.model.training:
            return None
        return self.rowgrad(self.table_for_target())

    def _early_link(self, rg, item_id, item_seq):
        """Both key sets of the step are known before the forward pass: thread them onto the per-row lists on a side stream while
        the encoder runs.  History keys go first, so uniq[0:n_hist) are the rows whose gradient needs the encoder backward and
        uniq[n_hist:n_uniq) are touched by the scorer only -- FusedOptimizer.early_apply updates the latter right after the
        loss kernel, concurrently with the backward pass."""
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
            self._ev_main = torch.cuda.Event()
            self._ev_link = torch.cuda.Event()
        main = torch.cuda.current_stream()
        self._ev_main.record(main)                      # the previous step's update of head/next/uniq is complete
        self._side.wait_event(self._ev_main)
        n_target, n_hist = item_id.numel(), item_seq.numel()
        with torch.cuda.stream(self._side):
            rg.prepare(n_target + n_hist)
            rg.n_uniq.zero_()
            ops.rowlist_link(rg.head, item_seq, n_target, rg.next, rg.uniq, rg.n_uniq, pad_id=rg.pad_id)
            rg.n_hist.copy_(rg.n_uniq)
            ops.rowlist_link(rg.head, item_id, 0, rg.next, rg.uniq, rg.n_uniq, pad_id=rg.pad_id)
            self._ev_link.record(self._side)
        rg.linked = True
        rg.early = True

    # ---- sum-pool tower hooks (overridden by the row-sharded engine) -----------------------------
    def pool_forward(self, item_seq, item_seq_len, user_id, use_seq, use_user, alpha):
        ws = self.ws
        d = self.table_for_target().shape[1]
        utable = self.model.user_embedding.weight if use_user else None
        if not use_seq:
            user = ws.get('user_emb', (user_id.shape[0], d))
            ops.gather_rows(utable.data, user_id, out=user)
            coeff = None
        else:
            B = item_seq.shape[0]
            user = ws.get('user_emb', (B, d))
            coeff = ws.get('pool_coeff', (B,))
            ops.pool_sum_fwd(self.table_for_seq().data, item_seq, item_seq_len, alpha,
                             utable.data if utable is not None else None, user_id, out=user, coeff_out=coeff)
        self._pool_saved = (item_seq, coeff, user_id)
        return user

    def pool_backward(self, d_user, use_seq, use_user):
        item_seq, coeff, user_id = self._pool_saved
        if use_seq:
            L = item_seq.shape[1]
            self.rowgrad(self.table_for_seq()).add(item_seq, d_user, L, coeff, L)
        if use_user:
            self.rowgrad(self.model.user_embedding.weight).add(user_id, d_user, 1, None, 1)

    # ---- sequence-row hooks (overridden by the row-sharded engine) -------------------------------
    def seq_rows_source(self, item_seq):
        """(table, index) such that table[index] are the history rows of the local batch."""
        return self.table_for_seq().data, item_seq

    def add_seq_rowgrad(self, item_seq, drows):
        self.rowgrad(self.table_for_seq()).add(item_seq, drows, 1, None, 1)

    # ---- forward / backward ---------------------------------------------------------------------
    def user_emb(self, save=False, **batch):
        self.ensure_ready()
        user = self.tower.forward(save=save, **batch)
        if not save:
            self.release_lo_plane()          # inference forward: no backward will close the lo-plane window
        return user

    def scores_only(self, user_emb, item_id, user_id=None):
        """_predict_layer without loss (eval / predict): recommender.py:76-96."""
        m, ws = self.model, self.ws
        item_id2 = item_id.view(item_id.shape[0], -1).contiguous()
        B, N = item_id2.shape
        scores = torch.empty(B, N, dtype=torch.float32, device=self.device)
        lv = ws.get('loss_vec_eval', (B,))
        # forward-only: softmax branch handles N == 1
        ops.score_loss(self.table_for_target().data, user_emb, item_id2, 'softmax',
                       item_bias=m.item_bias.data if m.has_item_bias else None,
                       user_bias=m.user_bias.data if m.has_user_bias else None, user_id=user_id, tau=m.tau,
                       score_clip=m.SCORE_CLIP, norm_host=1.0, scores=scores, loss_vec=lv)
        return scores.view(item_id.shape)

    # ---- evaluation helpers (overridden by the row-sharded engine) -------------------------------
    def eval_shard(self):
        """(world, rank) of the row partition the local target table holds."""
        return 1, 0

    def eval_sum(self, t):
        """Sum a per-user partial result over the ranks holding the table (identity for an unsharded table)."""
        return t

    def target_rows(self, idx):
        """Rows of the target table for global ids `idx` (forward_item_emb, recommender.py:66-74)."""
        return ops.gather_rows(self.table_for_target().data, idx)

    def rank_one_vs_all(self, user_emb, target, user_id=None, hist=None):
        """Number of catalogue items scoring above each user's target, history / padding / target excluded (csrc/evalrank.cu;
        reference: evaluator_abc.py:190-278 + onepos.py:20-31).  hist = (ptr int64, items, sorted int32) CSR device tensors."""
        self.ensure_ready()
        m, ws = self.model, self.ws
        B = user_emb.shape[0]
        W, r = self.eval_shard()
        table = self.table_for_target().data
        ib = m.item_bias.data if m.has_item_bias else None
        ub = m.user_bias.data if m.has_user_bias else None
        uid = user_id.contiguous() if user_id is not None else None
        if ub is not None and uid is None:
            raise ValueError('one_vs_all ranking with user_bias needs user_id')
        user_emb, target = user_emb.contiguous(), target.contiguous()
        tscore = ws.get('rank_tscore', (B,))
        counts = ws.get('rank_counts', (B,), dtype=torch.int32, zero=True)
        kw = dict(item_bias=ib, user_bias=ub, user_id=uid, tau=m.tau, world=W, rank=r)
        ops.rank_target(table, user_emb, target, tscore, **kw)
        self.eval_sum(tscore)
        ops.rank_count(table, user_emb, target, tscore, counts, **kw)
        ops.rank_exclude(table, user_emb, target, tscore, counts, hist_ptr=hist[0] if hist else None,
                         hist_sorted=hist[2] if hist else None, **kw)
        self.eval_sum(counts)
        return counts.clone()

    def forward_loss(self, user_id=None, item_id=None, label=None, item_seq=None, item_seq_len=None, reduction=True,
                     want_scores=False):
        self.ensure_ready()
        m, ws = self.model, self.ws
        if m.group_size > 0:
            item_id = item_id.view(-1, m.group_size)
            label = label.view(-1, m.group_size) if label is not None else None
        if item_id.dim() == 1:
            item_id = item_id.view(-1, 1)
            label = label.view(-1, 1) if label is not None else None
        item_id = item_id.contiguous()
        B, N = item_id.shape
        for rg in self._rowgrads.values():
            rg.reset()
        if self.uses_dropout and m.training:
            ops.rng_advance(self.rng)
        early_rg = self._overlap_rowgrad()
        if early_rg is not None:
            self._early_link(early_rg, item_id, item_seq)
        user = self.tower.forward(item_seq=item_seq, item_seq_len=item_seq_len, user_id=user_id, save=True)
        loss_type = m.loss_type
        scores = ws.get('scores', (B, N))
        loss_vec = ws.get('loss_vec', (B,))
        dscore = ws.get('dscore', (B, N))
        grad_user = ws.get('grad_user', (B, user.shape[1]))
        norm_dev, norm_host = None, float(B * max(N - 1, 1))
        if loss_type == 'softmax':
            if label is not None:
                label = label.contiguous()
                norm_dev = ws.get('n_pos', (1,))
                ops.count_positive(label, norm_dev)
            else:
                norm_host = float(B)
        ops.score_loss(self.table_for_target().data, user, item_id, loss_type, label=label if loss_type == 'softmax' else None,
                       item_bias=m.item_bias.data if m.has_item_bias else None,
                       user_bias=m.user_bias.data if m.has_user_bias else None, user_id=user_id, tau=m.tau,
                       score_clip=m.SCORE_CLIP, norm_dev=norm_dev, norm_host=norm_host, scores=scores, loss_vec=loss_vec,
                       dscore=dscore, grad_user=grad_user)
        loss = torch.empty((), dtype=torch.float32, device=self.device)
        if loss_type == 'softmax':
            ops.loss_finish(loss_vec, loss, denom_dev=norm_dev, denom_host=norm_host, nan_flag=self.nan_flag)
        else:
            ops.loss_finish(loss_vec, loss, denom_host=float(B), nan_flag=self.nan_flag)
        self.last = dict(user=user, item_id=item_id, user_id=user_id, dscore=dscore, grad_user=grad_user, B=B, N=N)
        if early_rg is not None and self.overlap_mode >= 2:
            self.overlap_hook.early_apply(self, early_rg, [(user, N, dscore, 1, B * N), (user, 1, None, 1, item_seq.numel())])
        out_loss = loss if reduction else loss_vec.clone()
        return out_loss, (scores if want_scores else None), user

    def backward(self, grad_out=None):
        """Backward of the last forward_loss.  Dense grads are ACCUMULATED into flat.grad; table grads are
        registered as row lists.  `grad_out` (dLoss upstream) is applied only when given."""
        st, m = self.last, self.model
        d_user, dscore = st['grad_user'], st['dscore']
        if grad_out is not None:
            d_user = d_user * grad_out
            dscore = dscore * grad_out
        fp = self.flat
        if m.has_item_bias:
            ops.scatter_add_scalar(fp.g('item_bias'), st['item_id'], dscore.contiguous())
        if m.has_user_bias:
            ops.scatter_add_scalar(fp.g('user_bias'), st['user_id'].contiguous(), dscore.contiguous(), idx_group=st['N'])
        self.rowgrad(self.table_for_target()).add(st['item_id'], st['user'], st['N'], dscore, 1)
        self.tower.backward(d_user)
        self.release_lo_plane()

    def zero_dense_grads(self):
        if self.flat is not None:
            self.flat.grad.zero_()
