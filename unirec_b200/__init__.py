"""unirec_b200 -- B200-native (sm_100a) hot path behind microsoft/UniRec's model / trainer / config API.

Package layout mirrors the reference where names are API (model registry walks file names):
  model/{base,sequential,cf}/   drop-in model classes (SASRec, GRU, AvgHist, SVDPlusPlus, MF)
  facility/                     Trainer, single-process/torch.distributed Accelerator, fused optimizer
  utils/, constants/, config/   argument parser, registry, YAML defaults (same keys and precedence)
  csrc/ + ../include            hand-written CUDA kernels behind a C ABI (ctypes-loaded, no fallback)
"""
__version__ = '0.1.0'


def install_as_unirec():
    """Register this package under the name `unirec` so code written against the reference
    (`from unirec.main import main`, `get_class_instance(name, 'unirec/model')`) resolves here."""
    import importlib
    import sys
    pkg = importlib.import_module(__name__)
    sys.modules.setdefault('unirec', pkg)
    for sub in ('model', 'model.base', 'model.sequential', 'model.cf', 'utils', 'constants', 'facility', 'main'):
        mod = importlib.import_module(__name__ + '.' + sub)
        sys.modules.setdefault('unirec.' + sub, mod)
    return pkg
