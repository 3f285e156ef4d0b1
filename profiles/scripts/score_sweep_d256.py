"""Variant sweep of the fused score/loss kernel at the c4 row width (d=256, N=4097): GB/s of table rows read."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from unirec_b200 import _cabi, ops
lib = _cabi.lib()
lib.ur_score_loss_set_bulk.argtypes = [ctypes.c_int]
V, d, B, N = 10_000_000, 256, 1024, 4097
table = torch.randn(V, d, device='cuda') * 0.02
g = torch.Generator(device='cuda').manual_seed(1)
ids = [torch.randint(1, V, (B, N), device='cuda', generator=g) for _ in range(4)]
lab = torch.zeros(B, N, dtype=torch.int32, device='cuda'); lab[:, 0] = 1
u = torch.randn(B, d, device='cuda') * 0.02
scores, dscore = torch.empty(B, N, device='cuda'), torch.empty(B, N, device='cuda')
lv, gu, npos = torch.empty(B, device='cuda'), torch.empty(B, d, device='cuda'), torch.full((1,), float(B), device='cuda')
ref = None
for mode, name in ((13, 'default 4l 6w 2st x2'), (20, '4l 6w 3st x1'), (21, '4l 4w 3st x2'), (22, '4l 3w 4st x2'), (23, '8l 6w 2st x2'), (24, '4l 8w 2st x1'), (25, '4l 4w 2st x3')):
    lib.ur_score_loss_set_bulk(mode)
    def run(i):
        ops.score_loss(table, u, ids[i % 4], 'softmax', label=lab, norm_dev=npos, scores=scores, loss_vec=lv, dscore=dscore, grad_user=gu)
    for i in range(3): run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): run(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    run(0); torch.cuda.synchronize()
    out = (lv.clone(), gu.clone())
    if ref is None: ref = out
    err = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(out, ref))
    print('%-22s %.4f ms  %.0f GB/s  (%.1f%% of 6539)  max rel diff %.2e' % (name, ms, B * N * d * 4 / ms / 1e6, B * N * d * 4 / ms / 1e6 / 65.395, err), flush=True)
lib.ur_score_loss_set_bulk(9)
