"""Minimal Accelerator: the surface of HF `accelerate.Accelerator` that the reference trainer/main use
(`.device .is_local_main_process .prepare .backward .clip_grad_norm_ .gather_for_metrics .unwrap_model
.wait_for_everyone .save`), implemented directly on torch.distributed -- one process per GPU, NCCL on CUDA
(gloo on CPU for the host-logic tests).  `accelerate` itself is not a dependency.
"""
import os

import torch
import torch.distributed as dist


class _PreparedOptimizer:
    """Mirrors accelerate's AcceleratedOptimizer: forwards everything, exposes the wrapped one as `.optimizer`
    (read at unirec/facility/trainer.py:396)."""

    def __init__(self, optimizer):
        self.optimizer = optimizer

    def __getattr__(self, name):
        return getattr(self.optimizer, name)

    def step(self, *a, **k):
        return self.optimizer.step(*a, **k)

    def zero_grad(self, *a, **k):
        return self.optimizer.zero_grad(*a, **k)

    def state_dict(self):
        return self.optimizer.state_dict()

    @property
    def param_groups(self):
        return self.optimizer.param_groups


class _DeviceLoader:
    """Wraps a DataLoader: every batch (tuple/list/dict of tensors) is moved to the device, non-blocking."""

    def __init__(self, loader, device):
        self.loader, self.device = loader, device
        self.dataset = getattr(loader, 'dataset', None)

    def __len__(self):
        return len(self.loader)

    def _move(self, x):
        if torch.is_tensor(x):
            return x.to(self.device, non_blocking=True)
        if isinstance(x, (list, tuple)):
            return type(x)(self._move(v) for v in x)
        if isinstance(x, dict):
            return {k: self._move(v) for k, v in x.items()}
        return x

    def __iter__(self):
        for batch in self.loader:
            yield self._move(batch)

    def __getattr__(self, name):
        return getattr(self.loader, name)


class Accelerator:
    def __init__(self, device=None, backend=None):
        self.num_processes = int(os.environ.get('WORLD_SIZE', '1'))
        self.process_index = int(os.environ.get('RANK', '0'))
        self.local_process_index = int(os.environ.get('LOCAL_RANK', '0'))
        if device is None:
            if torch.cuda.is_available():
                device = torch.device('cuda', self.local_process_index)
                torch.cuda.set_device(device)
            else:
                device = torch.device('cpu')
        self.device = torch.device(device)
        if self.num_processes > 1 and not dist.is_initialized():
            backend = backend or ('nccl' if self.device.type == 'cuda' else 'gloo')
            dist.init_process_group(backend=backend)
        self._optimizers = []

    @property
    def is_local_main_process(self):
        return self.local_process_index == 0

    @property
    def is_main_process(self):
        return self.process_index == 0

    @property
    def distributed(self):
        return self.num_processes > 1 and dist.is_initialized()

    def prepare(self, *objs):
        out = []
        for o in objs:
            if isinstance(o, torch.nn.Module):
                o = o.to(self.device)
                if hasattr(o, 'config'):
                    o.config['device'] = self.device
                    o.device = self.device
            elif isinstance(o, torch.optim.Optimizer):
                o = _PreparedOptimizer(o)
                self._optimizers.append(o)
            elif isinstance(o, torch.utils.data.DataLoader):
                o = _DeviceLoader(o, self.device)
            out.append(o)
        return out[0] if len(out) == 1 else tuple(out)

    def backward(self, loss):
        loss.backward()

    def clip_grad_norm_(self, parameters, max_norm, norm_type=2):
        """With a FusedOptimizer the clip is folded into the next `step()` (norm over flat grads + row lists, on
        device); otherwise torch's utility runs on the materialised gradients."""
        from .optim import FusedOptimizer
        fused = [o.optimizer for o in self._optimizers if isinstance(o.optimizer, FusedOptimizer)]
        if fused:
            for o in fused:
                o.max_grad_norm = float(max_norm)
            return None
        return torch.nn.utils.clip_grad_norm_(parameters, max_norm, norm_type)

    def gather_for_metrics(self, t):
        if not self.distributed:
            return t
        t = t.contiguous() if t.dim() else t.reshape(1)
        bufs = [torch.empty_like(t) for _ in range(self.num_processes)]
        dist.all_gather(bufs, t)
        return torch.cat(bufs, 0)

    def unwrap_model(self, model):
        return getattr(model, 'module', model)

    def wait_for_everyone(self):
        if self.distributed:
            dist.barrier()

    def save(self, obj, path):
        if self.is_main_process:
            torch.save(obj, path)

    def print(self, *a, **k):
        if self.is_local_main_process:
            print(*a, **k)


def broadcast(t, from_process=0):
    """accelerate.utils.broadcast as used at unirec/main/main.py:461-462."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(t, src=from_process)
    return t
