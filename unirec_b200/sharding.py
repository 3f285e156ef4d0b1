"""Row-sharded embedding tables across the GPUs of one box (replaces the reference's replicated tables +
DDP all-reduce of dense [V,d] gradients, unirec/facility/trainer.py:67,346).

Layout: rank r of W owns table rows {id : id % W == r}, stored at local index id // W (uniform NVSwitch fabric -> no
topology awareness needed; modulo spreads popular low ids).  The batch stays data-parallel (B samples per rank).

Per step (all collectives are fixed-size NCCL calls, no host synchronisation, the whole step is CUDA-graph capturable):
  ids      ur_pack_ids: (item id, label) -> one int32 per entry; all_gather (asynchronously, under the tower)  [S = W*B samples]
  history  SASRec: rows read from the owners' HBM over NVLink inside the gather+LayerNorm kernel (peer mappings, csrc/p2p.cu);
           GRU: owners gather into a zero-filled [S*L, d] buffer -> reduce_scatter; sum-pool towers (AvgHist / SVD++ / MF): owners
           pre-reduce their rows per sample (ur_pool_sum over owned rows) -> reduce_scatter of [S, d] partial user vectors
  tower    local (replicated encoder) -> u [B,d];  all_gather(u) -> U_all [S, d]
  scoring  "move queries, not rows".  softmax: ur_score_partial (cp.async.bulk ring over the OWNED target rows) -> per-sample
           online-softmax partial states; ONE all_to_all brings the W partials of a sample to its home rank, ur_score_merge ->
           loss, lse, dLoss/du; all_gather(lse); ur_score_dscore at the owner.
           bpr: owned raw scores -> reduce_scatter -> ur_bpr_from_scores at home -> all_gather(dLoss/ds) -> owners' partial
           dLoss/du -> reduce_scatter.
  backward tower backward -> row gradients stay where they are produced: owners pull them over NVLink inside the optimizer kernel
           (SASRec) or they are all-gathered (GRU / pool towers: [S, d] only for the pool towers)
  update   row-sparse optimizer on the local shards (ur_rowlist_link filters the gathered global ids by owner); encoder gradients
           and the reported loss: ONE all_reduce(SUM) of the flat gradient buffer (+1 slot for the loss).
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


# parameters stored row-sharded when table_shard_world > 1, and the attribute holding their full row count
SHARDED_TABLES = {'item_embedding.weight': 'n_items', 'item_dst_embedding.weight': 'n_items', 'item_src_embedding.weight': 'n_items',
                  'user_embedding.weight': 'n_users'}


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
    done = {}
    for name, rows_attr in SHARDED_TABLES.items():
        if name in sd:
            key = sd[name].data_ptr()                  # aliased names of one Parameter (item_src_embedding) are gathered once
            if key not in done:
                done[key] = gather_full_table(sd[name], int(getattr(model, rows_attr)), W, r, group)
            if r == 0:
                out[name] = done[key]
    return out


def localize_state_dict(model, state_dict):
    """Inverse for loading: a full [n_items, d] table in `state_dict` (a reference checkpoint, or one written by
    full_state_dict) is cut down to this rank's rows; already-local shards pass through."""
    W, r = int(getattr(model, 'shard_world', 1)), int(getattr(model, 'shard_rank', 0))
    if W <= 1:
        return state_dict
    out = dict(state_dict)
    for name, rows_attr in SHARDED_TABLES.items():
        t = out.get(name)
        if t is None:
            continue
        n = int(getattr(model, rows_attr))
        if t is not None and t.shape[0] == n and t.shape[0] != local_rows_count(n, W, r):
            out[name] = shard_table(t, W, r)
    return out


class ShardedEngine(Engine):
    """Engine whose embedding tables (item, item_dst, user) are row-sharded over `dist` ranks: every tower (SASRec, GRU, AvgHist,
    SVD++, MF) with softmax or BPR loss."""

    def __init__(self, model, tower_kind, world, rank, group=None):
        super().__init__(model, tower_kind)
        self.world, self.rank, self.group = int(world), int(rank), group
        # peer-memory mode (SASRec tower): history rows and their gradients are read straight from the owner / requester GPU over
        # NVLink (CUDA IPC mappings, csrc/p2p.cu) instead of a zero-filled [W, B*L, d] reduce-scatter and a [W, B*L, d] all-gather
        self.p2p = bool(int(model.config.get('shard_p2p', 1))) and tower_kind == 'sasrec' and self.world > 1
        self._peer_cache = {}
        self._prefetch = None           # packed ids whose all-gather is issued asynchronously under the tower
        self._pending = None

    # keys handed to the row lists are GLOBAL ids (packed or plain): ur_rowlist_link keeps the entries this rank owns
    def link_filter(self):
        return self.world, self.rank

    # ---- collectives --------------------------------------------------------------------------
    def _all_gather(self, name, t):
        out = self.ws.get('ag_' + name, (self.world,) + tuple(t.shape), dtype=t.dtype)
        if self.world == 1:
            out.copy_(t.view(out.shape))
            return out
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out

    def _all_gather_async(self, name, t):
        """all-gather on NCCL's stream, not waited for: (output, work, input kept alive)."""
        out = self.ws.get('ag_' + name, (self.world,) + tuple(t.shape), dtype=t.dtype)
        t = t.contiguous()
        if self.world == 1:
            out.copy_(t.view(out.shape))
            return out, None, t
        work = dist.all_gather_into_tensor(out, t, group=self.group, async_op=True)
        return out, work, t

    def _launch_prefetch(self):
        """The packed target ids (the bulk of the bytes exchanged per step) do not depend on the encoder: their all-gather starts
        right after the small history-id gather and runs on NCCL's stream while the tower computes."""
        if self._prefetch is None:
            return
        ids32 = self._prefetch
        self._prefetch = None
        self._pending = self._all_gather_async('ids32', ids32)

    def _reduce_scatter(self, name, t):
        """t: [W, ...] -> sum over ranks of slice [rank]."""
        out = self.ws.get('rs_' + name, tuple(t.shape[1:]), dtype=t.dtype)
        if self.world == 1:
            out.copy_(t[0])
            return out
        dist.reduce_scatter_tensor(out, t, op=dist.ReduceOp.SUM, group=self.group)
        return out

    def _all_to_all(self, name, t):
        """t: [W, ...] (slice w goes to rank w) -> [W, ...] (slice w came from rank w)."""
        if self.world == 1:
            return t
        out = self.ws.get('a2a_' + name, tuple(t.shape), dtype=t.dtype)
        dist.all_to_all_single(out, t, group=self.group)
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

    # ---- sequence rows (SASRec / GRU towers) --------------------------------------------------------
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
        keys = self.seq_all.view(-1)                                       # global ids of all ranks' histories
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

    # ---- sum-pool towers (AvgHist / SVD++ / MF): owners pre-reduce, partial user vectors are reduce-scattered --------------------
    def pool_forward(self, item_seq, item_seq_len, user_id, use_seq, use_user, alpha):
        ws, W, r = self.ws, self.world, self.rank
        d = self.table_for_target().shape[1]
        B = (item_seq if use_seq else user_id).shape[0]
        S = W * B
        utable = self.model.user_embedding.weight.data if use_user else None
        uid_all = self._all_gather('pool_uid', user_id.contiguous()).view(S) if use_user else None
        part = ws.get('pool_part', (W, B, d))
        if use_seq:
            seq_all = self._all_gather('seq', item_seq.contiguous())                      # [W, B, L]
            len_all = self._all_gather('pool_len', item_seq_len.contiguous()).view(S)
            coeff_all = ws.get('pool_coeff_all', (S,))
            ops.pool_sum_fwd(self.table_for_seq().data, seq_all.view(S, -1), len_all, alpha, utable, uid_all, out=part.view(S, d),
                             coeff_out=coeff_all, world=W, rank=r)
            self._pool_saved = (seq_all, coeff_all, uid_all)
        else:
            ops.shard_gather_rows(utable, uid_all, W, r, part.view(S, d))                 # zeros where not owned
            self._pool_saved = (None, None, uid_all)
        self._launch_prefetch()
        return self._reduce_scatter('pool_user', part)

    def pool_backward(self, d_user, use_seq, use_user):
        S = self.world * d_user.shape[0]
        d_all = self._all_gather('pool_duser', d_user).view(S, -1)         # dLoss/du of every sample: [S, d] only
        seq_all, coeff_all, uid_all = self._pool_saved
        if use_seq:
            L = seq_all.shape[-1]
            self.rowgrad(self.table_for_seq()).add(seq_all.view(-1), d_all, L, coeff_all, L)
        if use_user:
            self.rowgrad(self.model.user_embedding.weight).add(uid_all, d_all, 1, None, 1)

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
        ids32 = torch.empty(B, N, dtype=torch.int32, device=self.device)
        ops.pack_ids(item_id2, None, ids32)
        z = torch.empty(B, N, dtype=torch.float32, device=self.device)
        ops.shard_scores(self.table_for_target().data, user_emb.contiguous(), ids32, self.world, self.rank, z,
                         item_bias=m.item_bias.data if m.has_item_bias else None,
                         user_bias=m.user_bias.data if m.has_user_bias else None,
                         user_id=user_id if m.has_user_bias else None, tau=m.tau)
        self.eval_sum(z)
        if m.SCORE_CLIP > 0:
            z.clamp_(-m.SCORE_CLIP, m.SCORE_CLIP)
        return z.view(item_id.shape)

    # ---- forward / backward ---------------------------------------------------------------------
    def forward_loss(self, user_id=None, item_id=None, label=None, item_seq=None, item_seq_len=None, reduction=True,
                     want_scores=False):
        self.ensure_ready()
        m, ws, W, r = self.model, self.ws, self.world, self.rank
        if m.group_size > 0:
            item_id = item_id.view(-1, m.group_size)
            label = label.view(-1, m.group_size) if label is not None else None
        if item_id.dim() == 1:
            item_id = item_id.view(-1, 1)
            label = label.view(-1, 1) if label is not None else None
        item_id = item_id.contiguous()
        B, N = item_id.shape
        S = W * B
        for rg in self._rowgrads.values():
            rg.reset()
        if self.uses_dropout and m.training:
            ops.rng_advance(self.rng)
        loss_type = m.loss_type
        ids32 = ws.get('ids32', (B, N), dtype=torch.int32)
        ops.pack_ids(item_id, label.contiguous() if (label is not None and loss_type == 'softmax') else None, ids32)
        need_uid = m.has_user_bias and user_id is not None
        if m.has_user_bias and user_id is None:
            raise ValueError('has_user_bias needs user_id')
        self._prefetch, self._pending = ids32, None
        user = self.tower.forward(item_seq=item_seq, item_seq_len=item_seq_len, user_id=user_id, save=True)
        d = user.shape[1]
        self._launch_prefetch()                      # (towers that did not start it themselves)
        ids_out, work, _keep = self._pending
        self._pending = None
        if work is not None:
            work.wait()                              # the compute stream waits for NCCL's stream here
        ids_all = ids_out.view(S, N)
        uid_all = self._all_gather('uid', user_id.contiguous()).view(S) if need_uid else None
        u_all = self._all_gather('user', user).view(S, d)
        table = self.table_for_target().data
        ib = m.item_bias.data if m.has_item_bias else None
        ub = m.user_bias.data if m.has_user_bias else None
        loss_vec = ws.get('loss_vec', (B,))
        grad_user = ws.get('grad_user', (B, d))
        dscore = ws.get('dscore_all', (S, N))
        scores = None
        if loss_type == 'softmax':
            n_pos = ws.get('n_pos', (1,))
            ops.count_positive_packed(ids_all, n_pos)                        # global positives: same value on every rank
            z = ws.get('z_all', (S, N))
            if want_scores:
                z.zero_()
            state = ws.get('score_state', (W, B, 4 + 2 * d))
            ops.score_partial(table, u_all, ids_all, W, r, z, state.view(S, 4 + 2 * d), item_bias=ib, user_bias=ub, user_id=uid_all,
                              tau=m.tau, score_clip=m.SCORE_CLIP)
            mine = self._all_to_all('score_state', state)                    # [W, B, 4+2d]: partial of rank w for local sample b
            lse_ny = ws.get('lse_ny', (B, 2))
            ops.score_merge(mine, W, B, d, m.tau, n_pos, loss_vec, lse_ny, grad_user)
            lse_all = self._all_gather('lse', lse_ny).view(S, 2)
            ops.score_dscore(z, ids_all, lse_all, W, r, m.tau, m.SCORE_CLIP, n_pos, dscore)
            denom_dev, denom_host = n_pos, 1.0
            if want_scores:
                scores = self._reduce_scatter('scores', z.view(W, B, N))
                if m.SCORE_CLIP > 0:
                    scores = scores.clamp(-m.SCORE_CLIP, m.SCORE_CLIP)
        else:
            z = ws.get('z_all', (W, B, N))
            ops.shard_scores(table, u_all, ids_all, W, r, z.view(S, N), item_bias=ib, user_bias=ub, user_id=uid_all, tau=m.tau)
            z_home = self._reduce_scatter('z_home', z)                       # [B, N] complete raw scores of the local batch
            ds_home = ws.get('dscore_home', (B, N))
            sc_home = ws.get('scores_home', (B, N)) if want_scores else None
            ops.bpr_from_scores(z_home, m.tau, m.SCORE_CLIP, float(S * max(N - 1, 1)), loss_vec, ds_home, scores=sc_home)
            dscore = self._all_gather('dscore', ds_home).view(S, N)
            gpart = ws.get('grad_user_part', (W, B, d))
            ops.shard_grad_user(table, ids_all, dscore, W, r, gpart.view(S, d))
            grad_user = self._reduce_scatter('grad_user', gpart)
            denom_dev, denom_host = None, float(S)
            scores = sc_home
        # reported loss (not on the gradient path): local sum / global normaliser, summed over the ranks
        loss = torch.empty((), dtype=torch.float32, device=self.device)
        ops.loss_finish(loss_vec, loss, denom_dev=denom_dev, denom_host=denom_host, nan_flag=None)
        if W > 1:
            dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=self.group)
        ops.loss_finish(loss.view(1), loss, denom_host=1.0, nan_flag=self.nan_flag)       # NaN flag from the GLOBAL loss: all ranks agree
        self.last = dict(user=u_all, keys=ids_all.view(-1), dscore=dscore, grad_user=grad_user, B=B, N=N, S=S, ids_all=ids_all,
                         uid_all=uid_all)
        return (loss if reduction else loss_vec.clone()), scores, user

    def backward(self, grad_out=None):
        st, m = self.last, self.model
        d_user, dscore = st['grad_user'], st['dscore']
        if grad_out is not None:
            d_user = d_user * grad_out
            dscore = dscore * grad_out
        # bias gradients (replicated [V] / [n_users] vectors in the flat buffer): every rank adds the entries it owns (dscore is zero
        # elsewhere), the all-reduce of the flat gradient buffer completes the sum
        if m.has_item_bias:
            ops.scatter_add_scalar(self.flat.g('item_bias'), st['ids_all'], dscore.contiguous(), idx_mask=ops.PACKED_ID_MASK)
        if m.has_user_bias:
            ops.scatter_add_scalar(self.flat.g('user_bias'), st['uid_all'], dscore.contiguous(), idx_group=st['N'])
        self.rowgrad(self.table_for_target()).add(st['keys'], st['user'], st['N'], dscore, 1, key_mask=ops.PACKED_ID_MASK)
        self.tower.backward(d_user)
        self.release_lo_plane()

    def sync_dense_grads(self):
        """Encoder gradients add across ranks (global normalisation): ONE all-reduce of the flat buffer."""
        if self.flat is not None and self.flat.size and self.world > 1:
            dist.all_reduce(self.flat.grad, op=dist.ReduceOp.SUM, group=self.group)
