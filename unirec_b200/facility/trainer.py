"""Trainer: the SGD loop of the hot path (drop-in for unirec/facility/trainer.py:21-458, MoRec control excluded).

Control flow is the reference's, including its observable quirks: validation runs BEFORE each training
epoch, early stopping / best-checkpoint on the key metric, `optimizer.zero_grad()` after the forward, a NaN loss
skips the update, the logged epoch loss is the SUM of per-batch mean losses.  What changed underneath:
  * forward/backward/optimizer are the CUDA engine + FusedOptimizer (row-sparse table update, flat dense update);
  * no host synchronisation inside the step: the NaN guard is a device flag consumed by the optimizer kernels, the
    epoch loss is accumulated on the device and read once per epoch;
  * dense gradients of the (replicated) encoder are all-reduced as one flat buffer when running multi-process.
"""
import logging
import os
import time

import torch
import torch.distributed as dist
import torch.optim as optim

from unirec_b200.constants.protocols import DataFileFormat, EvaluationProtocal
from unirec_b200.utils.general import dict2str
from .evaluation import RankEvaluator
from .optim import FusedOptimizer


class Trainer(object):
    def __init__(self, config, model, accelerator):
        self.config = config
        self.exp_name = config['exp_name'] if 'exp_name' in config else __name__
        self.model = model
        self.accelerator = accelerator
        self.logger = logging.getLogger(self.exp_name)
        self.learning_rate = config.get('learning_rate', 0)
        self.epochs = config.get('epochs', 0)
        self.eval_step = min(1, self.epochs)
        self.early_stop = config.get('early_stop', 0)
        self.valid_metric_bigger = True
        self.test_batch_size = config.get('batch_size', 0)
        self.device = config.get('device', None)
        if 'checkpoint_dir' in config:
            self.checkpoint_dir = os.path.join(config['output_path'], config['checkpoint_dir'])
        else:
            self.checkpoint_dir = os.path.join(config['output_path'], 'checkpoint_{0}_{1}'.format(
                config.get('logger_time_str', 'run'), config.get('logger_rand', 0)))
        os.makedirs(self.checkpoint_dir, exist_ok=True)
        self.saved_model_file = os.path.join(self.checkpoint_dir, '{}.pth'.format(self.exp_name))
        self.weight_decay = config.get('weight_decay', 0)
        self.key_metric = config['key_metric'] if 'key_metric' in config else 'group_auc'
        self.best_valid_result = None
        self.best_valid_score = None
        self.start_epoch = 0
        self.cur_step = 1
        gcv = config.get('grad_clip_value', None)
        self.grad_clip_value = gcv if gcv is not None and gcv > 0 else None
        self.optimizer = self._build_optimizer(config['optimizer'], self.model)
        self.scheduler = self._build_scheduler(config['scheduler'], config['scheduler_factor'])
        self.model, self.optimizer, self.scheduler = self.accelerator.prepare(self.model, self.optimizer, self.scheduler)
        self._broadcast_replicated_params()
        self.evaluator = None
        self.user_history = None
        self.tb_logger = None

    def _broadcast_replicated_params(self):
        """Multi-process start: every replicated parameter takes rank 0's value (what DDP / accelerate.prepare do at
        unirec/facility/trainer.py:67).  Row-sharded tables are per-rank data and stay."""
        if not self.accelerator.distributed:
            return
        model = self.accelerator.unwrap_model(self.model)
        eng = getattr(model, '_engine', None)
        if eng is None or self.accelerator.device.type != 'cuda':
            return
        eng.ensure_ready()
        if eng.flat is not None and eng.flat.size:
            dist.broadcast(eng.flat.data, src=0)
        if int(getattr(model, 'shard_world', 1)) <= 1:
            for p in eng.table_params():
                dist.broadcast(p.data, src=0)

    # ------------------------------------------------------------------ factories
    def _build_optimizer(self, opt_type, model):
        """reference: trainer.py:134-152.  adam / adamw / sgd / sparse_adam run on the fused kernels; adagrad and
        rmsprop fall back to torch.optim over dense gradients (requires table_update=dense)."""
        if opt_type in ('adam', 'adamw', 'sgd', 'sparse_adam'):
            model._ur_fast_grads = True
            opt = FusedOptimizer(model, opt_type, lr=self.learning_rate, weight_decay=self.weight_decay)
            # overlap the row-sparse table update with the encoder backward (engine._early_link / optim.early_apply)
            mode = int(self.config.get('overlap_table_update', 0))
            if mode and not self.accelerator.distributed and hasattr(model, '_engine'):
                model._engine.overlap_hook = opt
                # mode 2 (early update of the target-only rows) needs the final gradient scale up front: no clipping
                model._engine.overlap_mode = mode if self.grad_clip_value is None else 1
            return opt
        if model.table_update != 'dense':
            raise ValueError("optimizer %r runs through torch.optim and needs dense table gradients: set table_update='dense'"
                             % (opt_type,))
        params = model.parameters()
        if opt_type == 'adagrad':
            return optim.Adagrad(params, lr=self.learning_rate, weight_decay=self.weight_decay)
        if opt_type == 'rmsprop':
            return optim.RMSprop(params, lr=self.learning_rate, weight_decay=self.weight_decay)
        self.logger.warning('Received unrecognized optimizer, set default Adam optimizer')
        model._ur_fast_grads = True
        return FusedOptimizer(model, 'adam', lr=self.learning_rate)

    def _build_scheduler(self, scheduler_type, factor):
        if scheduler_type == 'step':
            return optim.lr_scheduler.StepLR(self.optimizer, step_size=1, gamma=factor)
        if scheduler_type == 'reduce':
            return optim.lr_scheduler.ReduceLROnPlateau(self.optimizer, mode='max', factor=factor, patience=1, threshold=0.0001,
                                                        threshold_mode='rel', cooldown=0, min_lr=0, eps=1e-08)
        return None

    def set_user_history(self, user_history):
        self.user_history = user_history

    def reset_evaluator(self, data_format=None, eval_protocol=None):
        if eval_protocol in (EvaluationProtocal.OneVSAll.value, EvaluationProtocal.OneVSK.value) and \
                data_format not in (DataFileFormat.T5.value, DataFileFormat.T6.value):
            self.evaluator = RankEvaluator(self.config['metrics'], self.config.get('group_size', -1), self.config,
                                           self.accelerator, protocol=eval_protocol, user_history=self.user_history)
        else:
            raise ValueError('data format and evaluation protocol not match: {0} / {1}'.format(data_format, eval_protocol))

    @staticmethod
    def early_stopping(value, best, cur_step, max_step=4, bigger=True):
        """Validation-based early stopping (reference: trainer.py:186-233).  Returns (best, cur_step, stop, update)."""
        if max_step <= 0:
            return best, cur_step, False, True
        better = best is None or (value > best if bigger else value < best)
        if better:
            return value, 0, False, True
        cur_step += 1
        stop = cur_step > max_step if bigger else cur_step >= max_step
        return best, cur_step, stop, False

    # ------------------------------------------------------------------ training
    def _sync_dense_grads(self):
        """Multi-process: the encoder is replicated, its flat gradient buffer is averaged with ONE all-reduce
        (replaces DDP's bucketed all-reduce of every parameter incl. whole tables)."""
        if self.accelerator.distributed:
            eng = self.accelerator.unwrap_model(self.model)._engine
            if hasattr(eng, 'sync_dense_grads'):
                eng.sync_dense_grads()        # sharded tables: loss normalised globally, gradients add
                return
            flat = eng.flat
            if flat is not None and flat.size:
                dist.all_reduce(flat.grad, op=dist.ReduceOp.SUM)
                flat.grad.div_(self.accelerator.num_processes)

    def device_batches(self, batches, depth=2):
        """Iterate host batches (dicts or tuples of CPU tensors) as device batches, copying batch i+1 to the GPU on a copy
        stream while step i computes (pinned memory makes the copy asynchronous).  Device-resident batches pass through.
        The reference gets the same effect from accelerate's prepared DataLoader moving tensors ahead of the step
        (trainer.py:261)."""
        dev = self.accelerator.device
        if dev.type != 'cuda':
            yield from batches
            return
        copy_stream = getattr(self, '_copy_stream', None)
        if copy_stream is None:
            copy_stream = self._copy_stream = torch.cuda.Stream(device=dev)

        def to_dev(b):
            with torch.cuda.stream(copy_stream):
                if isinstance(b, dict):
                    out = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in b.items()}
                else:
                    out = type(b)(v.to(dev, non_blocking=True) if torch.is_tensor(v) else v for v in b)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return out, ev

        queue = []
        it = iter(batches)
        for b in it:
            queue.append(to_dev(b))
            if len(queue) >= depth:
                break
        while queue:
            out, ev = queue.pop(0)
            torch.cuda.current_stream().wait_event(ev)
            for v in (out.values() if isinstance(out, dict) else out):
                if torch.is_tensor(v) and v.is_cuda:
                    v.record_stream(torch.cuda.current_stream())
            nxt = next(it, None)
            if nxt is not None:
                queue.append(to_dev(nxt))
            yield out

    def train_step(self, samples):
        """One iteration of the hot loop (reference: trainer.py:327-357) without host synchronisation.  After two eager
        iterations per batch signature the whole step (forward, loss, backward, table + encoder update) is captured into a CUDA
        graph and replayed: ~45 kernel launches and the Python glue collapse into one graph launch."""
        g = self._graph_for(samples)
        if g is None:
            return self._eager_step(samples)
        for k, v in g['static'].items():
            v.copy_(samples[k], non_blocking=True)
        g['graph'].replay()
        return g['loss'].clone()

    # ---- CUDA-graph plumbing ------------------------------------------------------------------
    def _graph_for(self, samples):
        from unirec_b200 import ops
        from unirec_b200.engine import Engine
        model = self.accelerator.unwrap_model(self.model)
        eng = getattr(model, '_engine', None)
        from unirec_b200.sharding import ShardedEngine
        # the row-sharded step (NCCL collectives + peer-memory kernels) is captured too: `cuda_graph_sharded: 0` keeps it eager
        sharded_ok = type(eng) is ShardedEngine and int(self.config.get('cuda_graph_sharded', 1))
        if (not int(self.config.get('cuda_graph', 1)) or not (type(eng) is Engine or sharded_ok)
                or (self.accelerator.distributed and not sharded_ok)
                or not isinstance(self.optimizer, FusedOptimizer) and not isinstance(getattr(self.optimizer, 'optimizer', None), FusedOptimizer)
                or ops.PROFILE is not None or ops.TIMED_OP is not None or ops.TIMED_OPS is not None or eng.overlap_hook is not None):
            return None
        tensors = {k: v for k, v in samples.items() if torch.is_tensor(v)}
        if not tensors or any(not v.is_cuda for v in tensors.values()) or len(tensors) != len(samples):
            return None
        pg = self.optimizer.param_groups[0]
        key = (tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(tensors.items())),
               pg['lr'], pg['weight_decay'], tuple(pg['betas']), pg['eps'], self.grad_clip_value, model.training)
        cache = self.__dict__.setdefault('_graphs', {})
        g = cache.get(key)
        if g is None:
            if len(cache) >= 8:                 # a stream of ever-changing shapes (or learning rates) stays eager
                return None
            g = cache[key] = {'seen': 0}
        if 'graph' in g:
            return g
        g['seen'] += 1
        if g['seen'] <= 2:                      # eager warm-up: sizes the workspace, the row lists and the optimizer state
            return None
        static = {k: v.clone() for k, v in tensors.items()}
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss = self._eager_step(static)
        g.update(graph=graph, static=static, loss=loss)
        return g

    def _eager_step(self, samples):
        model = self.model
        model.train()
        loss, _, _, _ = model(**samples)
        self.optimizer.zero_grad()
        self.accelerator.backward(loss)
        self._sync_dense_grads()
        if self.grad_clip_value is not None:
            self.accelerator.clip_grad_norm_(model.parameters(), self.grad_clip_value)
        self.optimizer.step()          # skipped on-device when the loss is NaN
        return loss.detach()

    def fit(self, train_data, valid_data=None, save_model=True, load_pretrained_model=False, model_file=None, verbose=2):
        logger = self.logger
        if load_pretrained_model:
            if model_file is None:
                raise ValueError('`model_file` should be given when `load_pretrained_model` is set to True.')
            self.load_model(model_file)
        train_data, valid_data = self.accelerator.prepare(train_data, valid_data)
        key2index = train_data.dataset.return_key_2_index
        for epoch_idx in range(self.start_epoch, self.epochs):
            if valid_data is not None and self.evaluator is not None and (epoch_idx + 1) % self.eval_step == 0:
                t0 = time.time()
                valid_result = self.evaluate(valid_data, load_best_model=False, verbose=verbose)
                valid_score = valid_result[self.config['key_metric']]
                self.best_valid_score, self.cur_step, stop_flag, update_flag = Trainer.early_stopping(
                    valid_score, self.best_valid_score, self.cur_step, max_step=self.early_stop, bigger=self.valid_metric_bigger)
                logger.info('epoch %d evaluating [time: %.2fs, %s: %f]' % (epoch_idx, time.time() - t0, self.key_metric, valid_score))
                logger.info('complete scores on valid set: \n' + dict2str(valid_result))
                if update_flag:
                    if save_model:
                        self.accelerator.wait_for_everyone()
                        self.save_model(self.saved_model_file, self.optimizer, self.scheduler, epoch_idx, self.cur_step,
                                        valid_result, self.config)
                    self.best_valid_result = valid_result
                else:
                    logger.info('No better score in the epoch. Patience: {0} / {1}'.format(self.cur_step, self.early_stop))
                if stop_flag:
                    logger.info('Finished training, best eval result in epoch %d' % (epoch_idx - self.cur_step * self.eval_step))
                    break
                if self.scheduler and epoch_idx > 0:
                    if isinstance(self.scheduler, optim.lr_scheduler.ReduceLROnPlateau):
                        self.scheduler.step(valid_score)
                    else:
                        self.scheduler.step()
                    logger.info('epoch: %d, learning rate: %s' % (epoch_idx, self.optimizer.param_groups[0]['lr']))

            logger.info('\n>> epoch %d' % (epoch_idx + 1))
            t0 = time.time()
            total = None
            flag = self.accelerator.unwrap_model(self.model)._engine
            for batch_idx, inter_data in enumerate(self.device_batches(train_data)):
                samples = {k: inter_data[v] for k, v in key2index.items()}
                loss = self.train_step(samples)
                if self.accelerator.distributed and int(getattr(flag, 'world', 1)) <= 1:
                    loss = self.accelerator.gather_for_metrics(loss).mean()     # (the row-sharded engine reports the global loss)
                # NaN batches contribute nothing to the epoch loss (reference `continue`s before accumulating)
                contrib = torch.where(flag.nan_flag[0] != 0, torch.zeros_like(loss), loss)
                total = contrib if total is None else total + contrib
            total_loss = float(total) if total is not None else float('nan')      # one sync per epoch
            logger.info('epoch %d training [time: %.2fs, train loss: %.4f]' % (epoch_idx + 1, time.time() - t0, total_loss))
            self.last_epoch_loss = total_loss

    # ------------------------------------------------------------------ checkpoints
    def load_model(self, filename=None):
        checkpoint_file = filename if filename else self.saved_model_file
        checkpoint = torch.load(checkpoint_file, map_location=self.accelerator.device, weights_only=False)
        model = self.accelerator.unwrap_model(self.model)
        sd = checkpoint['state_dict']
        if int(getattr(model, 'shard_world', 1)) > 1:        # full tables in the file -> this rank's rows
            from unirec_b200.sharding import localize_state_dict
            sd = localize_state_dict(model, sd)
        model.load_state_dict(sd, strict=False)
        self.logger.info('Loading model from {0}. The best epoch was {1}'.format(checkpoint_file, checkpoint['cur_epoch']))
        if self.config.get('freeze', 0):
            for name, param in model.named_parameters():
                if name in checkpoint['state_dict']:
                    param.requires_grad = False

    def save_model(self, filename, optimizer, scheduler, cur_epoch=-1, cur_step=-1, best_valid_score=None, config=None):
        """Same checkpoint dict as the reference (trainer.py:389-398).  Row-sharded tables are re-assembled on rank 0 first
        (collective: every rank must call save_model), so the file holds full [n_items, d] tables like a reference checkpoint.
        Optimizer moments of sharded tables stay per-rank (saved by rank 0 for its shard only)."""
        model = self.accelerator.unwrap_model(self.model)
        if int(getattr(model, 'shard_world', 1)) > 1:
            from unirec_b200.sharding import full_state_dict
            model_sd = full_state_dict(model)
        else:
            model_sd = model.state_dict()
        state = {
            'config': config,
            'cur_epoch': cur_epoch,
            'cur_step': cur_step,
            'best_valid_score': best_valid_score,
            'state_dict': model_sd,
            'optimizer': optimizer.optimizer.state_dict() if optimizer is not None else None,
            'scheduler': scheduler.state_dict() if scheduler is not None else None,
        }
        for _ in range(5):
            try:
                self.accelerator.save(state, filename)
                self.logger.info('Saving best model at epoch {0} to {1}'.format(cur_epoch, filename))
                return
            except IOError:
                continue
        self.logger.error('Failed to save best model at epoch {0} to {1}'.format(cur_epoch, filename))

    # ------------------------------------------------------------------ evaluation
    @torch.no_grad()
    def evaluate(self, eval_data, load_best_model=True, model_file=None, verbose=0, predict_only=False):
        if load_best_model:
            self.load_model(model_file)
        if not hasattr(eval_data, 'device') and not isinstance(eval_data, type(None)):
            eval_data = self.accelerator.prepare(eval_data)
        model = self.accelerator.unwrap_model(self.model)
        return self.evaluator.evaluate(eval_data, model, verbose=verbose, predict_only=predict_only)
