"""BaseRecommender: forward orchestration behind the reference API (drop-in for
unirec/model/base/recommender.py:14-147), executed by the CUDA engine.

`forward(**batch)` keeps the reference signature and return convention:
  training -> (loss, None, None, None)   [or (loss, scores, user_emb, items_emb) with return_loss_only=False]
  eval     -> (None, scores, user_emb, items_emb)
The loss is a real autograd node: `loss.backward()` / `accelerator.backward(loss)` runs the engine's explicit
backward.  Dense (encoder) gradients reach `.grad` like any torch gradient; table gradients are
  * 'sparse' (default): kept as row lists on the engine and applied by unirec_b200.facility.optim.FusedOptimizer
                        (tables' `.grad` stays None);
  * 'dense'           : materialised as the reference's dense [V,d] gradient (config `table_update: dense`), so any
                        torch optimizer reproduces the reference step exactly.
"""
import inspect

import numpy as np
import torch

from unirec_b200.constants.loss_funcs import LossFuncType
from unirec_b200.engine import Engine
from .reco_abc import AbstractRecommender


class _LossNode(torch.autograd.Function):
    """Autograd anchor of the fused step: forward already happened in the engine; backward runs the engine's
    explicit backward and hands gradients to autograd for every parameter passed in."""

    @staticmethod
    def forward(ctx, loss_value, model, *params):
        ctx.model = model
        ctx.params = params
        ctx.n = len(params)
        return loss_value.clone()

    @staticmethod
    def backward(ctx, grad_out):
        model, eng = ctx.model, ctx.model._engine
        fast = getattr(model, '_ur_fast_grads', False)
        if fast:
            # trainer fast path: grads stay in the engine's flat buffer / row lists; upstream grad is 1
            eng.backward(None)
            return (None, None) + (None,) * ctx.n
        if grad_out.dim() > 0:
            raise RuntimeError('backward through reduction=False losses: reduce (e.g. .mean()) through the '
                               'engine instead; per-sample upstream gradients are not supported')
        eng.zero_dense_grads()
        eng.backward(grad_out)
        grads = []
        dense = {id(p): n for n, p in model.named_parameters() if n in eng.flat.offsets}
        rowgrads = {id(rg.param): rg for rg in eng.rowgrads()}
        for p in ctx.params:
            if id(p) in dense:
                grads.append(eng.flat.g(dense[id(p)]).clone())
            elif id(p) in rowgrads and model.table_update == 'dense':
                grads.append(rowgrads[id(p)].to_dense())
            else:
                grads.append(None)
        return (None, None) + tuple(grads)


class BaseRecommender(AbstractRecommender):
    _tower_kind = 'mf'

    def _init_attributes(self):
        super()._init_attributes()
        config = self.config
        self.dnn_inner_size = self.embedding_size
        self.time_seq = config.get('time_seq', 0)
        if self.time_seq:
            raise ValueError('time_seq embeddings are outside the accelerated hot path')
        self.table_update = str(config.get('table_update', 'sparse'))
        if self.table_update not in ('sparse', 'dense'):
            raise ValueError("table_update must be 'sparse' or 'dense'")

    def _init_modules(self):
        scorer_type = self.config['distance_type']
        if scorer_type != 'dot':
            if scorer_type in ('mlp', 'cosine'):
                raise ValueError("distance_type %r is outside the accelerated hot path; the fused scorer implements 'dot'"
                                 % scorer_type)
            raise ValueError('not supported distance_type: {0}'.format(scorer_type))
        super()._init_modules()
        if self.shard_world > 1 or self.config.get('table_shard_force', False):
            from unirec_b200.sharding import ShardedEngine
            self._engine = ShardedEngine(self, self._tower_kind, self.shard_world, self.shard_rank)
        else:
            self._engine = Engine(self, self._tower_kind)
        self._ur_fast_grads = False

    def _define_model_layers(self):
        pass

    # ---- towers ----------------------------------------------------------------------------------
    def _batch_for_tower(self, user_id, item_seq, item_seq_len):
        return dict(user_id=user_id, item_seq=item_seq, item_seq_len=item_seq_len)

    def forward_user_emb(self, user_id=None, item_seq=None, item_seq_len=None, item_seq_features=None, time_seq=None):
        with torch.no_grad():
            u = self._engine.user_emb(save=False, **self._batch_for_tower(user_id, item_seq, item_seq_len))
        return u.clone()

    def forward_item_emb(self, items, item_features=None):
        self._engine.ensure_ready()
        from unirec_b200 import ops
        idx = items if items.dtype in (torch.int32, torch.int64) else items.long()
        return self._engine.target_rows(idx.contiguous())          # row-sharded tables: owners contribute, ranks sum

    def item_embedding_for_user(self, item_seq, item_seq_features=None, time_seq=None):
        self._engine.ensure_ready()
        from unirec_b200 import ops
        return ops.gather_rows(self._engine.table_for_seq().data, item_seq.contiguous())

    def _predict_layer(self, user_emb, items_emb, user_id, item_id):
        """Scores for (user_emb, item ids).  The fused scorer reads the table rows itself, so `items_emb` is not
        consumed; it is accepted for signature compatibility (reference recommender.py:76-96)."""
        return self._engine.scores_only(user_emb.contiguous(), item_id, user_id)

    # ---- forward ---------------------------------------------------------------------------------
    def forward(self, user_id=None, item_id=None, label=None, item_features=None, item_seq=None, item_seq_len=None,
                item_seq_features=None, time_seq=None, session_id=None, reduction=True, return_loss_only=True, max_len=None):
        if self.loss_type == LossFuncType.FULLSOFTMAX.value:
            raise ValueError('fullsoftmax is outside the accelerated hot path')
        eng = self._engine
        if self.training:
            loss_raw, scores, user_emb = eng.forward_loss(user_id=user_id, item_id=item_id, label=label, item_seq=item_seq,
                                                          item_seq_len=item_seq_len, reduction=reduction,
                                                          want_scores=not return_loss_only)
            params = [p for p in self.parameters() if p.requires_grad]
            if not params:          # fully frozen model (load_model with `freeze`): keep the loss a differentiable node
                if getattr(self, '_frozen_anchor', None) is None or self._frozen_anchor.device != loss_raw.device:
                    self._frozen_anchor = torch.zeros((), device=loss_raw.device, requires_grad=True)
                params = [self._frozen_anchor]
            loss = _LossNode.apply(loss_raw, self, *params)
            if return_loss_only:
                return loss, None, None, None
            return loss, scores.clone().view(item_id.shape), user_emb.clone(), self.forward_item_emb(item_id)
        with torch.no_grad():
            user_emb = eng.user_emb(save=False, **self._batch_for_tower(user_id, item_seq, item_seq_len)).clone()
            scores = eng.scores_only(user_emb, item_id, user_id)
            items_emb = self.forward_item_emb(item_id)
        return None, scores, user_emb, items_emb

    def predict(self, interaction):
        inputs = {k: v for k, v in interaction.items() if k in inspect.signature(self.forward_user_emb).parameters}
        user_emb = self.forward_user_emb(**inputs)
        user_id = interaction['user_id'] if 'user_id' in interaction else None
        item_id = interaction['item_id'] if 'item_id' in interaction else None
        return self._predict_layer(user_emb, None, user_id, item_id).detach().cpu().numpy()

    def forward_all_item_emb(self, batch_size=None, numpy=True):
        """All item embeddings (reference recommender.py:108-128).  Without item features this is the table itself."""
        w = self._engine.table_for_target().data.detach()
        if self.shard_world > 1:        # collective: every rank re-assembles the full table in host memory
            from unirec_b200.sharding import gather_full_table_all
            full = gather_full_table_all(w, int(self.n_items), self.shard_world, self.shard_rank)
            return full.numpy().astype(np.float32) if numpy else full
        return w.cpu().numpy().astype(np.float32) if numpy else w.clone()

    def get_all_item_bias(self):
        return self.item_bias.detach().cpu().numpy()

    def get_user_bias(self, interaction):
        return self.user_bias[interaction['user_id']].detach().cpu().numpy()
