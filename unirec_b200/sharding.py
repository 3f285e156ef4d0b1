"""Row-sharded embedding tables across the GPUs of one box (replaces the reference's replicated tables +
DDP all-reduce of dense [V,d] gradients, unirec/facility/trainer.py:67,346).

Layout: rank r of W owns table rows {id : id % W == r}, stored at local index id // W (uniform NVSwitch fabric -> no
topology awareness needed; modulo spreads popular low ids).  The batch stays data-parallel (B samples per rank).

Per step (all collectives are fixed-size NCCL calls on torch's current stream, no host synchronisation):
  ids      all_gather(item_id, label, item_seq, user_id)                      -> every rank sees all W*B samples' ids
  history  owner gathers its rows into a zero-filled [W*B*L, d] buffer        -> reduce_scatter = local batch's rows
  tower    local (replicated encoder) -> u [B,d];  all_gather(u)              -> U_all [W*B, d]
  scoring  "move queries, not rows": ur_score_partial over OWNED target rows  -> per-sample online-softmax partials
           all_reduce(MAX) of the partial maxima, ur_score_rescale, reduce_scatter(SUM) of the partial states
           ur_score_finish at the home rank -> loss, lse, dLoss/du;  all_gather(lse)
           ur_score_dscore at the owner     -> per-entry dLoss/d(dot) for owned entries
  backward tower backward -> dX [B*L,d]; all_gather(dX); owners register (local row, source row, coef) lists
  update   row-sparse optimizer on the local shard (no exchange); encoder gradients: one all_reduce(SUM) of the flat
           buffer (the loss is normalised by the GLOBAL positive count, so gradients add across ranks).
Per-rank HBM traffic equals the single-GPU step (each rank reads 1/W of W batches' rows); NVLink carries ids, [W*B, O(d)]
states and the history rows.
"""
import torch
import torch.distributed as dist

from . import ops
from .engine import Engine


def owner_of(ids, world):
    return ids % world


def local_row(ids, world):
    return ids // world


def local_rows_count(n_rows, world, rank):
    """Rows of a [n_rows, d] table stored on `rank` (ids rank, rank+W, ...)."""
    return (n_rows - rank + world - 1) // world


def shard_table(full, world, rank):
    """Local shard of a full table (used by tests and checkpoint import)."""
    return full[rank::world].contiguous()


def unshard_tables(shards):
    """Inverse of shard_table over the list of per-rank shards."""
    world = len(shards)
    n = sum(s.shape[0] for s in shards)
    out = shards[0].new_empty((n,) + tuple(shards[0].shape[1:]))
    for r, s in enumerate(shards):
        out[r::world] = s
    return out


SHARDED_TABLES = ('item_embedding.weight',)       # parameters stored row-sharded when table_shard_world > 1


def gather_full_table_all(local, n_rows, world, rank, group=None, chunk_rows=1 << 20):
    """Every rank gets the full [n_rows, d] table in host memory (forward_all_item_emb of a sharded model,
    recommender.py:108-128): chunked all-gathers of equally sized (zero-padded) shard slices."""
    d = local.shape[1]
    full = torch.empty(n_rows, d, dtype=local.dtype)
    rows_max = local_rows_count(n_rows, world, 0)
    for c0 in range(0, rows_max, chunk_rows):
        c1 = min(rows_max, c0 + chunk_rows)
        mine = torch.zeros(c1 - c0, d, dtype=local.dtype, device=local.device)
        have = max(0, min(local.shape[0], c1) - c0)
        if have:
            mine[:have] = local[c0:c0 + have].detach()
        bufs = torch.empty(world, c1 - c0, d, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(bufs, mine, group=group)
        host = bufs.cpu()
        for src in range(world):
            n_src = max(0, min(local_rows_count(n_rows, world, src), c1) - c0)
            if n_src:
                full[src + c0 * world:src + (c0 + n_src - 1) * world + 1:world] = host[src, :n_src]
    return full


def gather_full_table(local, n_rows, world, rank, group=None, chunk_rows=1 << 20):
    """Re-assemble a row-sharded [n_rows, d] table on rank 0 (host memory) for a reference-compatible checkpoint
    (§8 f4; reference dict: unirec/facility/trainer.py:389-398).  Shards travel in chunks of `chunk_rows` rows, so the
    device-side staging stays small next to a 10M-row table.  Returns the full CPU tensor on rank 0, None elsewhere."""
    d = local.shape[1]
    full = torch.empty(n_rows, d, dtype=local.dtype) if rank == 0 else None
    for src in range(world):
        rows = local_rows_count(n_rows, world, src)
        for c0 in range(0, rows, chunk_rows):
            c1 = min(rows, c0 + chunk_rows)
            if src == 0:
                if rank == 0:
                    full[c0 * world:(c1 - 1) * world + 1:world] = local[c0:c1].detach().cpu()
                continue
            if rank == src:
                dist.send(local[c0:c1].detach().contiguous(), dst=0, group=group)
            elif rank == 0:
                buf = torch.empty(c1 - c0, d, dtype=local.dtype, device=local.device)
                dist.recv(buf, src=src, group=group)
                full[src + c0 * world:src + (c1 - 1) * world + 1:world] = buf.cpu()
    return full


def full_state_dict(model, group=None):
    """state_dict with every sharded table re-assembled (rank 0; other ranks get their local view back)."""
    sd = model.state_dict()
    W, r = int(getattr(model, 'shard_world', 1)), int(getattr(model, 'shard_rank', 0))
    if W <= 1:
        return sd
    out = dict(sd)
    for name in SHARDED_TABLES:
        if name in sd:
            full = gather_full_table(sd[name], int(model.n_items), W, r, group)
            if r == 0:
                out[name] = full
    return out


def localize_state_dict(model, state_dict):
    """Inverse for loading: a full [n_items, d] table in `state_dict` (a reference checkpoint, or one written by
    full_state_dict) is cut down to this rank's rows; already-local shards pass through."""
    W, r = int(getattr(model, 'shard_world', 1)), int(getattr(model, 'shard_rank', 0))
    if W <= 1:
        return state_dict
    out = dict(state_dict)
    for name in SHARDED_TABLES:
        t = out.get(name)
        if t is not None and t.shape[0] == int(model.n_items) and t.shape[0] != local_rows_count(int(model.n_items), W, r):
            out[name] = shard_table(t, W, r)
    return out


class ShardedEngine(Engine):
    """Engine whose embedding tables are row-sharded over `dist` ranks.  Softmax loss with SASRec / GRU towers (the
    north-star multi-GPU configuration); other combinations raise."""

    def __init__(self, model, tower_kind, world, rank, group=None):
        super().__init__(model, tower_kind)
        self.world, self.rank, self.group = int(world), int(rank), group
        if tower_kind not in ('sasrec', 'gru'):
            raise ValueError('row-sharded tables are implemented for the SASRec and GRU towers')
        if model.loss_type != 'softmax':
            raise ValueError('row-sharded tables are implemented for loss_type=softmax')
        # peer-memory mode (SASRec tower): history rows and their gradients are read straight from the owner / requester GPU over
        # NVLink (CUDA IPC mappings, csrc/p2p.cu) instead of a zero-filled [W, B*L, d] reduce-scatter and a [W, B*L, d] all-gather
        self.p2p = bool(int(model.config.get('shard_p2p', 1))) and tower_kind == 'sasrec' and self.world > 1
        self._peer_cache = {}
        self._prefetch = None           # (item_id, label, user_id) whose all-gathers are issued asynchronously under the tower
        self._pending = None

    def rowgrad(self, param):
        rg = super().rowgrad(param)
        rg.pad_id = -1             # keys are localized: -1 = not owned / global padding id
        return rg

    # ---- collectives --------------------------------------------------------------------------
    def _all_gather(self, name, t):
        out = self.ws.get('ag_' + name, (self.world,) + tuple(t.shape), dtype=t.dtype)
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out

    def _all_gather_async(self, name, t):
        """all-gather on NCCL's stream, not waited for: (output, work, input kept alive)."""
        out = self.ws.get('ag_' + name, (self.world,) + tuple(t.shape), dtype=t.dtype)
        t = t.contiguous()
        work = dist.all_gather_into_tensor(out, t, group=self.group, async_op=True)
        return out, work, t

    def _launch_prefetch(self):
        """The target / label id matrices (the bulk of the bytes all-gathered per step) do not depend on the encoder: their
        all-gathers start right after the small history-id gather and run on NCCL's stream while the tower computes."""
        if self._prefetch is None:
            return
        item_id, label, user_id = self._prefetch
        self._prefetch = None
        pend = {'ids': self._all_gather_async('ids', item_id), 'label': self._all_gather_async('label', label)}
        if user_id is not None:
            pend['uid'] = self._all_gather_async('uid', user_id)
        self._pending = pend

    def _reduce_scatter(self, name, t):
        """t: [W, ...] -> sum over ranks of slice [rank]."""
        out = self.ws.get('rs_' + name, tuple(t.shape[1:]), dtype=t.dtype)
        dist.reduce_scatter_tensor(out, t, op=dist.ReduceOp.SUM, group=self.group)
        return out

    # ---- peer memory ------------------------------------------------------------------------------
    def _peer_ptrs(self, name, t):
        """int64 device tensor [W]: pointers to every rank's buffer `name` (same shape everywhere), mapped through CUDA IPC.
        Collective on first use per buffer; the mappings live as long as the process."""
        key = (name, t.data_ptr(), tuple(t.shape))
        ptrs = self._peer_cache.get(key)
        if ptrs is None:
            handle, off = ops.ipc_export(t)
            gathered = [None] * self.world
            dist.all_gather_object(gathered, (handle, off), group=self.group)
            vals = [t.data_ptr() if w == self.rank else ops.ipc_open(h, o) for w, (h, o) in enumerate(gathered)]
            ptrs = torch.tensor(vals, dtype=torch.int64, device=self.device)
            self._peer_cache[key] = ptrs
        return ptrs

    # ---- sequence rows ----------------------------------------------------------------------------
    def seq_rows_source(self, item_seq):
        B, L = item_seq.shape
        d = self.table_for_seq().shape[1]
        seq_all = self._all_gather('seq', item_seq)                       # [W, B, L]
        self.seq_all = seq_all
        self._launch_prefetch()
        if self.p2p:
            # The all-gather above doubles as the step barrier: it completes only after every rank has enqueued it, i.e. after
            # every rank's previous table update -- the peers' shards are stable until their next optimizer step, which comes
            # after the all-reduce of the encoder gradients (= after every rank's backward pass has read them again).
            self.seq_shards = (self._peer_ptrs('table', self.table_for_seq().data), self.world)
            return self.table_for_seq().data, item_seq
        rows = self.ws.get('seq_rows_all', (self.world, B * L, d))
        ops.shard_gather_rows(self.table_for_seq().data, seq_all, self.world, self.rank, rows)
        mine = self._reduce_scatter('seq_rows', rows)                     # [B*L, d] rows of the local batch
        index = self.ws.get('seq_arange', (B, L), dtype=torch.int32)
        if getattr(self, '_arange_n', None) != B * L:
            index.copy_(torch.arange(B * L, dtype=torch.int32, device=self.device).view(B, L))
            self._arange_n = B * L
        return mine, index

    def add_seq_rowgrad(self, item_seq, drows):
        keys = self.ws.get('seq_keys_local', (self.seq_all.numel(),), dtype=torch.int32)
        ops.shard_localize(self.seq_all, self.world, self.rank, keys, pad_id=0)
        if self.p2p:
            # owners pull the gradient rows of their history entries from the requesters' `drows` buffers inside the optimizer
            # kernel (entry e of rank w's batch = row e of w's buffer).  Ordering: the optimizer runs after the all-reduce of the
            # encoder gradients (every rank has finished writing drows); a rank overwrites drows only after the next step's
            # first all-gather (every rank has finished reading it).
            parts = (self._peer_ptrs('drows', drows), drows.shape[0])
            self.rowgrad(self.table_for_seq()).add(keys, drows, 1, None, 1, parts=parts)
            return
        d_all = self._all_gather('drows', drows)                           # [W, B*L, d]
        self.rowgrad(self.table_for_seq()).add(keys, d_all.view(-1, d_all.shape[-1]), 1, None, 1)

    # ---- evaluation: every rank runs the SAME evaluation batches (the eval loaders are not sharded across ranks), each rank works on
    # the rows it owns and the partial results add up -- no table rows move ------------------------------------------------------
    def eval_shard(self):
        return self.world, self.rank

    def eval_sum(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def target_rows(self, idx):
        idx = idx.contiguous()
        out = torch.empty(tuple(idx.shape) + (self.table_for_target().shape[1],), dtype=torch.float32, device=self.device)
        ops.shard_gather_rows(self.table_for_target().data, idx, self.world, self.rank, out)     # zeros where not owned
        return self.eval_sum(out)

    def scores_only(self, user_emb, item_id, user_id=None):
        """_predict_layer without loss (eval / predict, recommender.py:76-96): owners score their entries, the rest is zero, sum."""
        m, ws = self.model, self.ws
        item_id2 = item_id.view(item_id.shape[0], -1).contiguous()
        B, N = item_id2.shape
        d = user_emb.shape[1]
        z = torch.zeros(B, N, dtype=torch.float32, device=self.device)
        state = ws.get('score_state_eval', (B, 4 + 2 * d))
        ops.score_partial(self.table_for_target().data, user_emb.contiguous(), item_id2, self.world, self.rank, z, state,
                          item_bias=m.item_bias.data if m.has_item_bias else None,
                          user_bias=m.user_bias.data if m.has_user_bias else None,
                          user_id=user_id if m.has_user_bias else None, tau=m.tau, score_clip=m.SCORE_CLIP)
        self.eval_sum(z)
        if m.SCORE_CLIP > 0:
            z.clamp_(-m.SCORE_CLIP, m.SCORE_CLIP)
        return z.view(item_id.shape)

    # ---- forward / backward ---------------------------------------------------------------------
    def forward_loss(self, user_id=None, item_id=None, label=None, item_seq=None, item_seq_len=None, reduction=True,
                     want_scores=False):
        self.ensure_ready()
        m, ws, W, r = self.model, self.ws, self.world, self.rank
        if item_id.dim() == 1:
            item_id = item_id.view(-1, 1)
            label = label.view(-1, 1) if label is not None else None
        item_id = item_id.contiguous()
        B, N = item_id.shape
        for rg in self._rowgrads.values():
            rg.reset()
        if self.uses_dropout and m.training:
            ops.rng_advance(self.rng)
        if label is None:
            label = ws.get('default_label', (B, N), dtype=torch.int32, zero=True)
            label[:, 0] = 1
        need_uid = m.has_user_bias and user_id is not None
        self._prefetch, self._pending = (item_id, label.contiguous(), user_id if need_uid else None), None
        user = self.tower.forward(item_seq=item_seq, item_seq_len=item_seq_len, user_id=user_id, save=True)
        d = user.shape[1]
        S = W * B
        self._launch_prefetch()                      # (towers that did not go through seq_rows_source)
        pend, self._pending = self._pending, None
        for out, work, _keep in pend.values():
            work.wait()                              # the compute stream waits for NCCL's stream here
        ids_all = pend['ids'][0].view(S, N)
        lab_all = pend['label'][0].view(S, N)
        uid_all = pend['uid'][0].view(S) if need_uid else None
        u_all = self._all_gather('user', user).view(S, d)
        n_pos = ws.get('n_pos', (1,))
        ops.count_positive(lab_all, n_pos)                                  # global positives: same value on every rank
        z = ws.get('z_all', (S, N))
        state = ws.get('score_state', (S, 4 + 2 * d))
        ops.score_partial(self.table_for_target().data, u_all, ids_all, W, r, z, state, label=lab_all,
                          item_bias=m.item_bias.data if m.has_item_bias else None,
                          user_bias=m.user_bias.data if m.has_user_bias else None, user_id=uid_all, tau=m.tau,
                          score_clip=m.SCORE_CLIP)
        gmax = ws.get('score_gmax', (S,))
        gmax.copy_(state[:, 0])
        dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=self.group)
        ops.score_rescale(state, gmax, d)
        mine = self._reduce_scatter('score_state', state.view(W, B, 4 + 2 * d))          # [B, 4+2d]
        loss_vec = ws.get('loss_vec', (B,))
        lse_ny = ws.get('lse_ny', (B, 2))
        grad_user = ws.get('grad_user', (B, d))
        ops.score_finish(mine, gmax.view(W, B)[r], d, m.tau, n_pos, loss_vec, lse_ny, grad_user)
        lse_all = self._all_gather('lse', lse_ny).view(S, 2)
        dscore = ws.get('dscore_all', (S, N))
        ops.score_dscore(z, ids_all, lab_all, lse_all, W, r, m.tau, m.SCORE_CLIP, n_pos, dscore)
        # loss: local sum / global positives, then summed over ranks = global mean (reported value, not on the grad path)
        loss = torch.empty((), dtype=torch.float32, device=self.device)
        ops.loss_finish(loss_vec, loss, denom_dev=n_pos, nan_flag=None)
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=self.group)
        self.nan_flag.copy_(torch.isnan(loss).to(torch.int32).view(1))
        keys = ws.get('tgt_keys_local', (S * N,), dtype=torch.int32)
        ops.shard_localize(ids_all, W, r, keys, pad_id=0)
        self.last = dict(user=u_all, keys=keys, dscore=dscore, grad_user=grad_user, B=B, N=N, ids_all=ids_all, S=S)
        scores = None
        if want_scores:
            zs = torch.where(ids_all % W == r, z, torch.zeros_like(z)).view(W, B, N)
            scores = self._reduce_scatter('scores', zs.contiguous())
            if m.SCORE_CLIP > 0:
                scores = scores.clamp(-m.SCORE_CLIP, m.SCORE_CLIP)
        return (loss if reduction else loss_vec.clone()), scores, user

    def backward(self, grad_out=None):
        st, m = self.last, self.model
        d_user, dscore = st['grad_user'], st['dscore']
        if grad_out is not None:
            d_user = d_user * grad_out
            dscore = dscore * grad_out
        if m.has_item_bias:
            self.flat.g('item_bias').index_add_(0, st['ids_all'].reshape(-1), dscore.reshape(-1))
        if m.has_user_bias:
            raise NotImplementedError('user_bias gradients with row-sharded tables')
        self.rowgrad(self.table_for_target()).add(st['keys'], st['user'], st['N'], dscore, 1)
        self.tower.backward(d_user)

    def sync_dense_grads(self):
        """Encoder gradients add across ranks (global normalisation): ONE all-reduce of the flat buffer."""
        if self.flat is not None and self.flat.size:
            dist.all_reduce(self.flat.grad, op=dist.ReduceOp.SUM, group=self.group)
