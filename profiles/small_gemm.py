"""B-row products of the trimmed last encoder layer (B = 1024, d = 128, inner 512): time per launch inside a CUDA graph
(no host launch cost), default route (ur_gemm_f32 / ur_gemm_fused_f32 dispatch).  Usage: python profiles/small_gemm.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unirec_b200 import ops

dev = torch.device('cuda', 0)
B, d, I = 1024, 128, 512
r = lambda *s: torch.randn(*s, device=dev)
x, w, w1, w2, b1 = r(B, d), r(d, d), r(I, d), r(d, I), r(I)
h, hp, y, dy, dh, gw1, gw2, gw = r(B, I), r(B, I), r(B, d), r(B, d), r(B, I), r(I, d), r(d, I), r(d, d)
shapes = [
    ('fwd  q/dense NT 1024x128x128', lambda P: ops.gemm(x, w, y, B, d, d, transB=True, precision=P)),
    ('fwd  ffn1    NT 1024x512x128 +act', lambda P: ops.gemm(x, w1, h, B, I, d, transB=True, bias=b1, act='swish', preact=hp, precision=P)),
    ('fwd  ffn2    NT 1024x128x512', lambda P: ops.gemm(h, w2, y, B, d, I, transB=True, precision=P)),
    ('dx   ffn2    NN 1024x512x128 *act\'', lambda P: ops.gemm_fused(dy, w2, dh, B, I, d, precision=P, dact=hp, act='swish', colsum=b1)),
    ('dx   ffn1    NN 1024x128x512 acc', lambda P: ops.gemm(dh, w1, dy, B, d, I, accumulate=True, precision=P)),
    ('dx   dense   NN 1024x128x128', lambda P: ops.gemm(dy, w, y, B, d, d, precision=P)),
    ('dW   ffn2    TN 128x512x1024', lambda P: ops.gemm(dy, h, gw2, d, I, B, transA=True, lda=d, accumulate=True, precision=P)),
    ('dW   ffn1    TN 512x128x1024', lambda P: ops.gemm(dh, x, gw1, I, d, B, transA=True, lda=I, accumulate=True, precision=P)),
    ('dW   dense   TN 128x128x1024', lambda P: ops.gemm(dy, x, gw, d, d, B, transA=True, lda=d, accumulate=True, precision=P)),
]
tot = 0.0
for name, fn in shapes:
    fn(3); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): fn(3)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 100
    tot += us
    print('%-40s %6.1f us per launch' % (name, us))
print('sum %.1f us' % tot)
