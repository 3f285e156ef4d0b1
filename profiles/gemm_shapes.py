"""Per-shape timing of the encoder GEMMs of the bench workload (T = 51200 tokens, d = 128, inner 512), CUDA events,
L2 flushed between launches.  Usage: python profiles/gemm_shapes.py  (prints a table; run under gpurun)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unirec_b200 import ops  # noqa: E402

dev = torch.device('cuda', 0)
T, d, I = 51200, 128, 512
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def run(name, fn, bytes_algo, reps=5):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    t = sorted(ts)[len(ts) // 2]
    return t, bytes_algo / t / 1e3


def main():
    g = torch.Generator(device=dev).manual_seed(0)
    r = lambda *s: torch.randn(*s, device=dev, generator=g)
    x, ctx, hact, hpre = r(T, d), r(T, d), r(T, I), r(T, I)
    wqkv, wo, w1, w2 = r(3 * d, d), r(d, d), r(I, d), r(d, I)
    bqkv, bo, b1, b2 = r(3 * d), r(d), r(I), r(d)
    qkv, z, hp, ha, dh, dx = r(T, 3 * d), r(T, d), r(T, I), r(T, I), r(T, I), r(T, d)
    gw_qkv, gw_o, gw1, gw2, gb1 = r(3 * d, d), r(d, d), r(I, d), r(d, I), r(I)
    MB = 1e6
    shapes = [
        ('fwd qkv   NT 51200x384x128', lambda P: ops.gemm(x, wqkv, qkv, T, 3 * d, d, transB=True, bias=bqkv, precision=P), (T * d + T * 3 * d) * 4),
        ('fwd dense NT 51200x128x128', lambda P: ops.gemm(ctx, wo, z, T, d, d, transB=True, bias=bo, precision=P), 2 * T * d * 4),
        ('fwd ffn1  NT 51200x512x128 +act+preact', lambda P: ops.gemm(x, w1, ha, T, I, d, transB=True, bias=b1, act='swish', preact=hp, precision=P), (T * d + 2 * T * I) * 4),
        ('fwd ffn2  NT 51200x128x512', lambda P: ops.gemm(hact, w2, z, T, d, I, transB=True, bias=b2, precision=P), (T * I + T * d) * 4),
        ('dx  ffn2  NN 51200x512x128 *act\' +colsum', lambda P: ops.gemm_fused(dx, w2, dh, T, I, d, precision=P, dact=hpre, act='swish', colsum=gb1), (T * d + 2 * T * I) * 4),
        ('dx  ffn1  NN 51200x128x512 accumulate', lambda P: ops.gemm(dh, w1, dx, T, d, I, accumulate=True, precision=P), (T * I + 2 * T * d) * 4),
        ('dx  dense NN 51200x128x128', lambda P: ops.gemm(dx, wo, z, T, d, d, precision=P), 2 * T * d * 4),
        ('dx  qkv   NN 51200x128x384 accumulate', lambda P: ops.gemm(qkv, wqkv, dx, T, d, 3 * d, accumulate=True, precision=P), (T * 3 * d + 2 * T * d) * 4),
        ('dW  ffn2  TN 128x512x51200', lambda P: ops.gemm(dx, hact, gw2, d, I, T, transA=True, lda=d, accumulate=True, precision=P), (T * d + T * I) * 4),
        ('dW  ffn1  TN 512x128x51200', lambda P: ops.gemm(dh, x, gw1, I, d, T, transA=True, lda=I, accumulate=True, precision=P), (T * d + T * I) * 4),
        ('dW  dense TN 128x128x51200', lambda P: ops.gemm(dx, ctx, gw_o, d, d, T, transA=True, lda=d, accumulate=True, precision=P), 2 * T * d * 4),
        ('dW  qkv   TN 384x128x51200', lambda P: ops.gemm(qkv, x, gw_qkv, 3 * d, d, T, transA=True, lda=3 * d, accumulate=True, precision=P), (T * 3 * d + T * d) * 4),
    ]
    if len(sys.argv) > 1:                       # substring filter, e.g. "ffn" (used for ncu captures)
        shapes = [sh for sh in shapes if any(a in sh[0] for a in sys.argv[1:])]
    tot = {1: 0.0, 3: 0.0}
    print('%-46s %9s %9s %9s %9s %8s' % ('shape', 'tf32 us', 'GB/s', '3xtf32 us', 'GB/s', 'MB algo'))
    for name, fn, nbytes in shapes:
        row = []
        for P in (1, 3):
            fn(P)
            torch.cuda.synchronize()
            t, gbs = run(name, lambda: fn(P), nbytes)
            tot[P] += t
            row += [t, gbs]
        print('%-46s %9.1f %9.0f %9.1f %9.0f %8.1f' % (name, row[0], row[1], row[2], row[3], nbytes / MB))
    print('sum per layer: tf32 %.1f us, 3xtf32 %.1f us' % (tot[1], tot[3]))


if __name__ == '__main__':
    main()
