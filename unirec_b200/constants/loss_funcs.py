from enum import Enum


class LossFuncType(Enum):
    """String values are config API (reference: unirec/constants/loss_funcs.py:6-11)."""
    BCE = 'bce'
    BPR = 'bpr'
    SOFTMAX = 'softmax'
    CCL = 'ccl'
    FULLSOFTMAX = 'fullsoftmax'


class DistanceType(Enum):
    DOT = 'dot'
    COSINE = 'cosine'
    MLP = 'mlp'


# losses implemented by the fused CUDA scorer (north-star scope)
SUPPORTED_LOSSES = (LossFuncType.SOFTMAX.value, LossFuncType.BPR.value)
