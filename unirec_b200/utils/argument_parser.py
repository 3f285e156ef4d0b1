"""Configuration assembly (drop-in for unirec/utils/argument_parser.py:214-241).

Result: one flat dict.  Precedence, lowest to highest:
    config/base.yaml < config/model/<model>.yaml < config/dataset/<dataset>.yaml < --config_file < command line < `args` dict
`--config_dir` overrides the YAML root; unknown command-line flags are ignored; flags given as none/None are dropped;
`config['cmd_args']` keeps the command-line + args layer so it can be re-applied over a checkpoint's config.
"""
import argparse
import os
from typing import Dict

from . import file_io

# flag name -> type.  Names and types follow the reference CLI (they are API); flags of out-of-scope subsystems
# (MoRec, VAE/SLIM solvers, ConvFormer, text/feature embeddings) are accepted so existing command lines keep parsing.
_FLAGS = {
    str: ['exp_name', 'config_file', 'config_dir', 'model', 'dataloader', 'task', 'wandb_file', 'checkpoint_dir', 'metrics',
          'key_metric', 'device', 'init_method', 'scheduler', 'optimizer', 'dataset', 'dataset_path', 'data_train_name',
          'data_valid_name', 'data_test_name', 'output_path', 'train_file_format', 'valid_file_format', 'test_file_format',
          'user_history_file_format', 'user_history_filename', 'test_protocol', 'valid_protocol', 'model_file',
          'features_filepath', 'features_shape', 'text_emb_path', 'item_emb_path', 'history_mask_mode', 'edge_norm',
          'loss_type', 'distance_type', 'hidden_act', 'padding_mode', 'linear_mode', 'train_type', 'base_model',
          'morec_objectives', 'morec_objective_controller', 'morec_objective_weights', 'item_meta_morec_filename',
          'align_dist_filename',
          # unirec_b200 additions
          'table_update', 'gemm_precision'],
    int: ['seed', 'use_wandb', 'use_tensorboard', 'num_workers', 'num_workers_train', 'num_workers_valid', 'num_workers_test',
          'epochs', 'gpu_id', 'verbose', 'shuffle_train', 'early_stop', 'batch_size', 'train_batch_size', 'valid_batch_size',
          'test_batch_size', 'n_sample_neg_train', 'n_sample_neg_valid', 'n_sample_neg_test', 'group_size',
          'load_pretrained_model', 'time_seq', 'seq_last', 'use_features', 'use_text_emb', 'text_emb_size', 'embedding_size',
          'use_pre_item_emb', 'use_position_emb', 'max_seq_len', 'has_user_bias', 'has_item_bias', 'hidden_size', 'inner_size',
          'n_layers', 'asymmetric', 'conv_size', 'seq_merge', 'encoder_dims', 'decoder_dims', 'total_anneal_steps',
          'eval_reparameter_sampling_times', 'freeze', 'enable_morec', 'morec_ngroup',
          'n_heads', 'n_users', 'n_items', 'has_user_emb',
          # unirec_b200 additions
          'pack_sequences', 'trim_last_layer', 'cuda_graph', 'overlap_table_update', 'shard_p2p', 'table_shard_world',
          'table_shard_rank'],
    float: ['grad_clip_value', 'score_clip_value', 'init_std', 'init_mean', 'scheduler_factor', 'learning_rate',
            'neg_by_pop_alpha', 'dropout_prob', 'hidden_dropout_prob', 'attn_dropout_prob', 'ccl_w', 'ccl_m', 'weight_decay',
            'layer_norm_eps', 'tau', 'user_sequence_alpha', 'seq_decay', 'init_ratio', 'anneal_cap', 'l1_coef', 'l2_coef',
            'item_spec_reg', 'admm_penalty', 'morec_alpha', 'morec_lambda', 'morec_expect_loss', 'morec_beta_min',
            'morec_beta_max', 'morec_K_p', 'morec_K_i'],
}


def parse_cmd_arguments(argv=None) -> Dict:
    parser = argparse.ArgumentParser(add_help=False)
    for typ, names in _FLAGS.items():
        for n in names:
            parser.add_argument('--' + n, type=typ)
    ns, _unknown = parser.parse_known_args(argv)
    return {k: v for k, v in sorted(vars(ns).items()) if v is not None and v not in ('none', 'None')}


def default_config_dir():
    return os.path.abspath(os.path.join(os.path.dirname(__file__), '..', 'config'))


def parse_arguments(args: Dict = None, argv=None) -> Dict:
    top = parse_cmd_arguments(argv)
    if args is not None:
        top.update(args)
    config_dir = top.get('config_dir', default_config_dir())
    print('Load configuration files from {}'.format(config_dir))
    config = file_io.load_yaml(os.path.join(config_dir, 'base.yaml'))
    config.update(file_io.load_yaml(os.path.join(config_dir, 'model', top['model'] + '.yaml')))
    config.update(file_io.load_yaml(os.path.join(config_dir, 'dataset', top['dataset'] + '.yaml')))
    if 'config_file' in top:
        print('Update configuration with the file {}'.format(top['config_file']))
        config.update(file_io.load_yaml(top['config_file']))
    config.update(top)
    config['cmd_args'] = top
    return config
