"""Entrypoint (drop-in for unirec/main/main.py:425-505): `run(args)` / `python -m unirec_b200.main.main --flag=value ...`.

config -> Accelerator -> user history -> model (registry) -> loaders -> Trainer.fit -> test evaluation -> result file.
What differs from the reference underneath the same contract:
  * training batches come from the device-side batch builder (`device_batch_builder: 1`, default on CUDA) instead of DataLoader
    worker processes; evaluation batches use the per-sample CPU datasets;
  * `accelerate`, wandb, tensorboard, MoRec and cProfile are not part of the hot path and are not wired in.
"""
import copy
import logging
import os
import random
import sys
import time
from typing import Dict

import numpy as np
import torch
from torch.utils.data import DataLoader

from unirec_b200.constants.protocols import DataFileFormat, EvaluationProtocal, TaskType
from unirec_b200.facility.accelerator import Accelerator, broadcast
from unirec_b200.facility.trainer import Trainer
from unirec_b200.utils import argument_parser, general


def get_user_history(config):
    """reference: main.py:106-118"""
    if config.get('user_history_filename') is None:
        return None, None
    return general.load_user_history(config['dataset_path'], config['user_history_filename'], config['n_users'],
                                     config.get('user_history_file_format', 'user-item'), config.get('time_seq', 0))


def construct_item_popularity(config, dataset_cls):
    """Item frequency in the training file (reference: main.py:235-245); feeds popularity^alpha negative sampling."""
    cfg = copy.deepcopy(config)
    cfg.update(data_loader_task='train', data_format=config['train_file_format'], eval_protocol=None)
    ds = dataset_cls(cfg, config['dataset_path'], config.get('data_train_name', 'train'))
    pop = np.bincount(ds.item_id, minlength=config['n_items']).astype(np.float64)
    pop[0] = 0
    return pop


def get_data_loader(config, task, add_history_trans, dataset_cls, file_path, file_name, user2history=None, item2popularity=None,
                    accelerator=None):
    """reference: main.py:121-204.  task in {'train','valid','test'}."""
    config = copy.deepcopy(config)
    config['data_loader_task'] = task
    config['data_format'] = config['{0}_file_format'.format(task)]
    config['eval_protocol'] = config.get('{0}_protocol'.format(task), None) if task != 'train' else None
    if config['eval_protocol'] == EvaluationProtocal.OneVSAll.value:
        config['n_sample_neg_{0}'.format(task)] = -1
    n_neg = config.get('n_sample_neg_{0}'.format(task), -1) or -1
    if '{0}_batch_size'.format(task) in config:
        config['batch_size'] = config['{0}_batch_size'.format(task)]
    on_device = (task == 'train' and n_neg > 0 and int(config.get('device_batch_builder', 1)) and accelerator is not None
                 and accelerator.device.type == 'cuda' and config['data_format'] in (DataFileFormat.T1.value, DataFileFormat.T3.value))
    transform = None
    if n_neg > 0 and not on_device:
        transform = general.get_class_instance('AddNegSamples', 'unirec_b200/data')(
            config['n_users'], config['n_items'], n_neg, user2history=user2history, item_popularity=item2popularity,
            neg_by_pop_alpha=config.get('neg_by_pop_alpha', None))
    dataset = dataset_cls(config, path=file_path, filename=file_name, transform=transform)
    if add_history_trans is not None:
        dataset.add_user_history_transform(add_history_trans(user2history, config['history_mask_mode'], None,
                                                              config.get('seq_last', 0), config['data_format']))
    if on_device:
        from unirec_b200.data.device_loader import DeviceBatchLoader
        return DeviceBatchLoader(dataset, config['batch_size'], accelerator.device, n_neg, config['n_users'], config['n_items'],
                                 max_seq_len=config.get('max_seq_len', 0) if add_history_trans is not None else 0,
                                 user_history=user2history, history_mask_mode=config['history_mask_mode'],
                                 seq_last=config.get('seq_last', 0), item_popularity=item2popularity,
                                 neg_by_pop_alpha=config.get('neg_by_pop_alpha', None), shuffle=bool(config['shuffle_train']),
                                 seed=config.get('seed', 2022), rank=accelerator.process_index, world=accelerator.num_processes)
    workers = config.get('num_workers_{0}'.format(task), config['num_workers'])
    return DataLoader(dataset, batch_size=config['batch_size'], shuffle=bool(config['shuffle_train']) if task == 'train' else False,
                      num_workers=workers, pin_memory=bool(config.get('pin_memory', False)),
                      persistent_workers=bool(config.get('persistent_workers', False)) and workers > 0)


def main(config, accelerator):
    logger = logging.getLogger(config['exp_name'])
    dataset_cls = general.get_class_instance(config['dataloader'], 'unirec_b200/data/dataset')
    is_seq = config['dataloader'] == 'SeqRecDataset'
    user2history, _ = get_user_history(config)
    add_hist = general.get_class_instance('AddUserHistory', 'unirec_b200/data') if is_seq else None
    item2pop = construct_item_popularity(config, dataset_cls) if config.get('neg_by_pop_alpha') else None

    model = general.get_class_instance(config['model'], 'unirec_b200/model')(config)
    task = config.get('task', TaskType.TRAIN.value)
    trainer = Trainer(config, model, accelerator)
    trainer.set_user_history(user2history)
    path = config['dataset_path']
    result = None
    if task == TaskType.TRAIN.value:
        train_data = get_data_loader(config, 'train', add_hist, dataset_cls, path, config.get('data_train_name', 'train'),
                                     user2history, item2pop, accelerator)
        valid_data = get_data_loader(config, 'valid', add_hist, dataset_cls, path, config.get('data_valid_name', 'valid'),
                                     user2history, item2pop, accelerator)
        trainer.reset_evaluator(config['valid_file_format'], config['valid_protocol'])
        try:
            trainer.fit(train_data, valid_data, save_model=bool(config.get('saved', True)),
                        load_pretrained_model=bool(config.get('load_pretrained_model', False)),
                        model_file=config.get('model_file', None), verbose=config.get('verbose', 2))
        except KeyboardInterrupt:
            logger.info('Keyboard interrupt: stop training and evaluate the best checkpoint')
        result = trainer.best_valid_result
    test_base = os.path.join(path, config.get('data_test_name', 'test'))
    have_test = any(os.path.exists(test_base + ext) for ext in ('.pkl', '.ftr', '.tsv', '.csv', '.txt'))
    if task in (TaskType.TRAIN.value, TaskType.TEST.value) and have_test:
        test_data = get_data_loader(config, 'test', add_hist, dataset_cls, path, config.get('data_test_name', 'test'),
                                    user2history, item2pop, accelerator)
        trainer.reset_evaluator(config['test_file_format'], config['test_protocol'])
        model_file = config.get('model_file', None) if task == TaskType.TEST.value else None
        have_ckpt = model_file is not None or os.path.exists(trainer.saved_model_file)
        result = trainer.evaluate(test_data, load_best_model=have_ckpt, model_file=model_file)
        logger.info('test result: {0}'.format(result))
        if accelerator.is_local_main_process:
            with open(os.path.join(config['output_path'], 'result_{0}.tsv'.format(config['exp_name'])), 'w') as f:
                f.write('\t'.join(sorted(result)) + '\n' + '\t'.join(str(result[k]) for k in sorted(result)) + '\n')
    return result


def run(args: Dict = None):
    config = argument_parser.parse_arguments(args)
    config.setdefault('exp_name', 'unirec_b200')
    config.setdefault('output_path', os.path.join(os.getcwd(), 'output'))
    config.setdefault('task', TaskType.TRAIN.value)
    for split in ('train', 'valid', 'test'):
        config.setdefault('{0}_file_format'.format(split), config.get('train_file_format', 'user-item'))
    os.makedirs(config['output_path'], exist_ok=True)
    if config.get('gpu_id', 0) is not None and int(config.get('gpu_id', 0)) >= 0 and 'WORLD_SIZE' not in os.environ:
        os.environ.setdefault('CUDA_VISIBLE_DEVICES', str(config['gpu_id']))
    accelerator = Accelerator()
    config['device'] = accelerator.device
    if accelerator.num_processes > 1:
        config.setdefault('table_shard_world', accelerator.num_processes)
        config.setdefault('table_shard_rank', accelerator.process_index)
    stamp = torch.tensor([int(time.time()), random.randint(0, 100000)], dtype=torch.int64, device=accelerator.device)
    stamp = broadcast(stamp, 0)
    config['logger_time_str'] = time.strftime('%Y-%m-%d_%H%M%S', time.localtime(int(stamp[0])))
    config['logger_rand'] = int(stamp[1])
    logger = logging.getLogger(config['exp_name'])
    if accelerator.is_local_main_process and not logger.handlers:
        logger.setLevel(logging.INFO)
        logger.addHandler(logging.StreamHandler(sys.stderr))
        logger.addHandler(logging.FileHandler(os.path.join(config['output_path'], '{0}.log'.format(config['exp_name']))))
    general.init_seed(config['seed'])
    return main(config, accelerator)


if __name__ == '__main__':
    print(run())
