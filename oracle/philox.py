"""CPU restatement of the dropout mask generator of csrc/dropout.cuh (TEST INFRASTRUCTURE ONLY).

Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; the Random123
library), restated in numpy from the published round function:

    (c0, c1, c2, c3), (k0, k1)  ->  (mulhi(M1, c2) ^ c1 ^ k0,  mullo(M1, c2),  mulhi(M0, c0) ^ c3 ^ k1,  mullo(M0, c0))
    key += (0x9E3779B9, 0xBB67AE85) after every round,  M0 = 0xD2511F53, M1 = 0xCD9E8D57, 10 rounds.

Pinned by the Random123 known-answer vectors (tests/test_dropout_oracle.py).  The reference (nn.Dropout at
unirec/model/sequential/sasrec.py:69, unirec/model/modules.py:307,313,352, unirec/model/sequential/gru.py:29) draws
its masks from torch's global generator, which cannot be reproduced by a fused kernel; parity with dropout is
therefore defined on EXPLICIT masks: the same multipliers are injected into the reference modules
(oracle/make_golden.py) and into this oracle.

Mask of element e of site s in training step t:  word (e % 4) of philox(counter = (e//4 lo, e//4 hi, t, s),
key = (seed lo, seed hi));  multiplier = 1/(1-p) if word >= floor(p * 2^32) else 0.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy arrays of counters (uint64 holding 32-bit words); k0, k1 Python ints."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK32 for c in (c0, c1, c2, c3))
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _threshold(p):
    t = float(p) * 4294967296.0
    return 0xFFFFFFFF if t >= 4294967295.0 else int(t)


def mask_flat(seed, step, site, p, n, start=0):
    """Multipliers of elements start .. start+n-1 of a site, float32 [n]."""
    if p <= 0.0:
        return np.ones(n, dtype=np.float32)
    e = np.arange(start, start + n, dtype=np.uint64)
    g = e >> np.uint64(2)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    words = philox4x32_10(g & MASK32, g >> np.uint64(32), np.full_like(g, int(step) & 0xFFFFFFFF), np.full_like(g, int(site)),
                          seed & 0xFFFFFFFF, seed >> 32)
    sel = (e & np.uint64(3)).astype(np.int64)
    w = np.choose(sel, words)
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return np.where(w >= np.uint64(_threshold(p)), scale, np.float32(0.0)).astype(np.float32)


def mask_rows(seed, step, site, p, rows, d):
    """[rows, d] multipliers of a site indexed by (token position, column): element position * d + column."""
    return mask_flat(seed, step, site, p, rows * d).reshape(rows, d)


def sasrec_masks(cfg, B, L, seed, step):
    """Every nn.Dropout of the SASRec tower as explicit multipliers (site numbering of unirec_b200/engine.py):
    'input' [B,L,d]; per layer i 'attn.i' [B,H,L,L], 'attn_out.i' [B,L,d], 'ffn_out.i' [B,L,d]."""
    import torch
    d, H = int(cfg['embedding_size']), int(cfg['n_heads'])
    ph, pa = float(cfg.get('hidden_dropout_prob', 0) or 0), float(cfg.get('attn_dropout_prob', 0) or 0)
    out = {'input': torch.from_numpy(mask_rows(seed, step, 0, ph, B * L, d)).view(B, L, d)}
    for i in range(int(cfg['n_layers'])):
        out['attn.%d' % i] = torch.from_numpy(mask_flat(seed, step, 1 + 3 * i, pa, B * H * L * L)).view(B, H, L, L)
        out['attn_out.%d' % i] = torch.from_numpy(mask_rows(seed, step, 2 + 3 * i, ph, B * L, d)).view(B, L, d)
        out['ffn_out.%d' % i] = torch.from_numpy(mask_rows(seed, step, 3 + 3 * i, ph, B * L, d)).view(B, L, d)
    return out


def gru_masks(cfg, B, L, seed, step):
    import torch
    d = int(cfg['embedding_size'])
    return {'input': torch.from_numpy(mask_rows(seed, step, 0, float(cfg.get('dropout_prob', 0) or 0), B * L, d)).view(B, L, d)}
