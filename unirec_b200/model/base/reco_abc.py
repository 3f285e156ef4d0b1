"""AbstractRecommender: attributes, tables, initialisation (drop-in for unirec/model/base/reco_abc.py).

Parameter creation stays in torch, in the reference's order, so seeds reproduce the reference's initial weights
bit for bit; everything numerical afterwards runs in the CUDA engine.
"""
import logging
import os

import torch
import torch.nn as nn
from torch.nn.init import constant_, xavier_normal_, xavier_uniform_

from unirec_b200.constants.loss_funcs import LossFuncType, SUPPORTED_LOSSES
from unirec_b200.constants.protocols import DataFileFormat


def _make_init(kind, mean=0.0, std=0.02):
    """Per-module initialiser (reference: reco_abc.py:19-57): Embedding/Linear weights drawn by `kind`,
    padding row and biases zero, LayerNorm = (1, 0)."""
    def draw(w):
        if kind == 'normal':
            w.normal_(mean=mean, std=std)
        elif kind == 'xavier_normal':
            xavier_normal_(w)
        elif kind == 'xavier_uniform':
            xavier_uniform_(w)
        else:
            raise KeyError(kind)

    def init(module):
        if isinstance(module, nn.Embedding):
            draw(module.weight.data)
            if module.padding_idx is not None:
                constant_(module.weight.data[module.padding_idx], 0.)
        elif isinstance(module, nn.Linear):
            draw(module.weight.data)
            if module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
    return init


class AbstractRecommender(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.logger = logging.getLogger(config['exp_name'])
        self.__optimized_by_SGD__ = True
        self.config = config
        self._init_attributes()
        self._init_modules()
        self.annotations = []
        self.add_annotation()
        self._parameter_validity_check()

    # ---- hooks for subclasses -------------------------------------------------------------------
    def _define_model_layers(self):
        raise NotImplementedError

    def forward(self, user_id):
        raise NotImplementedError

    def forward_user_emb(self, interaction):
        raise NotImplementedError

    def forward_item_emb(self, interaction):
        raise NotImplementedError

    def _predict_layer(self, user_emb, items_emb, interaction):
        raise NotImplementedError

    def predict(self, interaction):
        raise NotImplementedError

    def add_annotation(self):
        self.annotations.append('AbstractRecommender')

    # ---- shared behaviour -----------------------------------------------------------------------
    def _parameter_validity_check(self):
        if self.loss_type not in SUPPORTED_LOSSES:
            raise ValueError('unirec_b200 implements loss_type in %s on the fused CUDA scorer; got %r '
                             '(bce/ccl/fullsoftmax are outside the accelerated path)' % (SUPPORTED_LOSSES, self.loss_type))
        if self.loss_type == LossFuncType.SOFTMAX.value:
            if self.config['train_file_format'] in (DataFileFormat.T2.value, DataFileFormat.T2_1.value) and self.group_size <= 0:
                raise ValueError('SOFTMAX loss on user-item-label data needs a positive group_size '
                                 '(each positive line followed by the same number of negative lines).')

    def _init_attributes(self):
        config = self.config
        self.n_users = config['n_users']
        self.n_items = config['n_items']
        self.device = config['device']
        self.loss_type = config.get('loss_type', 'bce')
        self.embedding_size = config.get('embedding_size', 0)
        self.hidden_size = config.get('hidden_size', self.embedding_size)
        self.dropout_prob = config.get('dropout_prob', 0.0)
        self.init_method = config.get('init_method', 'normal')
        for flag in ('use_pre_item_emb', 'use_text_emb', 'use_features'):
            if config.get(flag, 0):
                raise ValueError('%s is outside the accelerated hot path (north-star scope) and is not supported' % flag)
        self.use_pre_item_emb = self.use_text_emb = self.use_features = 0
        self.group_size = config['group_size'] if 'group_size' in config else -1
        self.SCORE_CLIP = config['score_clip_value'] if 'score_clip_value' in config else -1
        self.has_user_bias = bool(config.get('has_user_bias', False))
        self.has_item_bias = bool(config.get('has_item_bias', False))
        self.tau = config.get('tau', 1.0)
        # unirec_b200 addition: row-sharding of the item tables over the ranks of one box (see unirec_b200/sharding.py)
        self.shard_world = int(config.get('table_shard_world', 1) or 1)
        self.shard_rank = int(config.get('table_shard_rank', os.environ.get('RANK', 0))) if self.shard_world > 1 else 0

    def _init_modules(self):
        # creation order is part of the RNG contract (reference reco_abc.py:159-189)
        if self.has_user_bias:
            self.user_bias = nn.Parameter(torch.normal(0, 0.1, size=(self.n_users,)))
        if self.has_item_bias:
            self.item_bias = nn.Parameter(torch.normal(0, 0.1, size=(self.n_items,)))
        # `table_init_device: 1` (unirec_b200 addition, used by bench.py): the big tables are created and initialised directly in
        # HBM (a 50M x 128 table never exists in host memory); values then come from the device generator, not the reference's
        tdev = self.device if int(self.config.get('table_init_device', 0) or 0) and torch.device(self.device).type == 'cuda' else None
        # row-sharded tables (multi-GPU): this rank stores rows {id : id % W == r}; the padding row lives on rank 0
        W, r = self.shard_world, self.shard_rank
        self._chunked_table_init = (W > 1 or tdev is not None) and self.init_method == 'normal'
        self._table_rows = {}            # id(weight) -> full row count of a table created through _make_table

        def _make_table(n_rows):
            rows = (n_rows - r + W - 1) // W
            if self._chunked_table_init:
                # no constructor draw for the shard; in host mode burn the n_rows x d N(0,1) draws nn.Embedding's own reset_parameters
                # makes in the unsharded model, so that the generator stays aligned with it (see _init_params)
                emb = nn.Embedding(rows, self.embedding_size, padding_idx=0 if r == 0 else None,
                                   _weight=torch.empty(rows, self.embedding_size, device=tdev))
                if tdev is None:
                    for c0 in range(0, n_rows, 1 << 18):
                        torch.empty(min(n_rows, c0 + (1 << 18)) - c0, self.embedding_size).normal_()
            else:
                emb = nn.Embedding(rows, self.embedding_size, padding_idx=0 if r == 0 else None, device=tdev)
            emb._ur_full_rows = n_rows
            return emb

        if self.config['has_user_emb']:
            self.user_embedding = _make_table(self.n_users)
        self.item_embedding = _make_table(self.n_items)
        self._define_model_layers()
        self._init_params()

    def _init_params(self):
        mean, std = self.config.get('init_mean', 0.0), self.config.get('init_std', 0.02)
        init = _make_init(self.init_method, mean, std)
        W, r = self.shard_world, self.shard_rank

        def init_sharded_table(emb):
            """Row shard of the table the UNSHARDED model would draw: the full [n_rows, d] stream is generated in row chunks and
            rows id % W == r are kept, so every rank consumes the same amount of the generator (the replicated encoder that
            follows is initialised identically on all ranks) and the W shards together equal the single-GPU initial table."""
            w = emb.weight.data
            n_rows = emb._ur_full_rows
            chunk = 1 << 18
            for c0 in range(0, n_rows, chunk):
                c1 = min(n_rows, c0 + chunk)
                full = torch.empty(c1 - c0, w.shape[1], dtype=w.dtype, device=w.device).normal_(mean=mean, std=std)
                first = (r - c0) % W                      # first row of the chunk owned by this rank
                mine = full[first::W]
                lo = (c0 + first) // W
                w[lo:lo + mine.shape[0]] = mine
            if r == 0:
                w[0] = 0.0

        done = set()
        for _, module in self.named_children():
            if self._chunked_table_init and isinstance(module, nn.Embedding) and hasattr(module, '_ur_full_rows'):
                if id(module.weight) not in done:
                    done.add(id(module.weight))
                    init_sharded_table(module)
                continue
            module.apply(init)
