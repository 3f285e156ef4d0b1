"""Alias-method sampling tables (Walker/Vose) for popularity^alpha negative sampling.
API counterpart of unirec/utils/sampling.py:9-31 (`prepare_aliased_randomizer`); the tables are also uploaded to the GPU for
the device-side batch builder (csrc/batch.cu)."""
import numpy as np


def build_alias_table(weights):
    """weights >= 0 (need not be normalised).  Returns (prob float32 [n], alias int32 [n]): draw slot s uniformly, keep s with
    probability prob[s], else take alias[s]."""
    w = np.asarray(weights, dtype=np.float64)
    n = w.shape[0]
    scaled = w * (n / w.sum())
    prob = np.ones(n, dtype=np.float64)
    alias = np.arange(n, dtype=np.int32)
    small = [i for i in range(n) if scaled[i] < 1.0]
    large = [i for i in range(n) if scaled[i] >= 1.0]
    while small and large:
        s, l = small.pop(), large.pop()
        prob[s], alias[s] = scaled[s], l
        scaled[l] = scaled[l] + scaled[s] - 1.0
        (small if scaled[l] < 1.0 else large).append(l)
    for i in small + large:      # numerical leftovers
        prob[i] = 1.0
    return prob.astype(np.float32), alias


def prepare_aliased_randomizer(weights, rng=None):
    prob, alias = build_alias_table(weights)
    rng = rng or np.random.default_rng()
    n = len(prob)

    def draw():
        s = int(rng.integers(n))
        return s if rng.random() < prob[s] else int(alias[s])
    return draw


def popularity_weights(item_popularity, alpha):
    """popularity^alpha with the padding id excluded (reference: addnegsamples.py:58-62)."""
    w = np.power(np.asarray(item_popularity, dtype=np.float64), alpha)
    w[0] = 0.0
    return w / w.sum()
