"""Micro-benchmark of the fused score/loss kernel variants (A/B switch ur_score_loss_set_bulk): GB/s of table rows read."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unirec_b200 import _cabi, ops
lib = _cabi.lib()
lib.ur_score_loss_set_bulk.argtypes = [ctypes.c_int]
V, d, B = 10_000_000, 128, 1024
table = torch.randn(V, d, device='cuda') * 0.02
for N in (1025,):
    g = torch.Generator(device='cuda').manual_seed(1)
    ids = [torch.randint(1, V, (B, N), device='cuda', generator=g) for _ in range(8)]
    lab = torch.zeros(B, N, dtype=torch.int32, device='cuda'); lab[:, 0] = 1
    u = torch.randn(B, d, device='cuda') * 0.02
    scores, dscore = torch.empty(B, N, device='cuda'), torch.empty(B, N, device='cuda')
    lv, gu, npos = torch.empty(B, device='cuda'), torch.empty(B, d, device='cuda'), torch.full((1,), float(B), device='cuda')
    ref = None
    for mode, name in ((0, 'register-staged v1'), (2, 'bulk ring v2'), (10, 'v3 2 lanes/row, 4 warps'), (11, 'v3 4 lanes/row, 8 warps'), (12, 'v3 2l 3w 4st'), (13, 'v3 2l 6w 2st'), (14, 'v3 4l 2w 2st x7'), (15, 'v3 2l 1w 2st x7'), (16, 'v3 4l 2w 2st x4'), (17, 'v3 2l 2w 2st x5'), (18, 'v3 2l 12w 2st x1'), (19, 'v3 2l 8w 3st x1'), (26, 'v3 2l 10w 2st x1'), (27, 'v3 2l 8w 2st x1')):
        lib.ur_score_loss_set_bulk(mode)
        def run(i):
            ops.score_loss(table, u, ids[i % 8], 'softmax', label=lab, norm_dev=npos, scores=scores, loss_vec=lv, dscore=dscore, grad_user=gu)
        for i in range(5): run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(40): run(i)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 40
        run(0); torch.cuda.synchronize()
        out = (lv.clone(), gu.clone(), dscore.clone())
        if ref is None: ref = out
        err = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(out, ref))
        print('N=%d %-22s %.4f ms  %.0f GB/s  (%.1f%% of 6553)  max rel diff vs v1 %.2e' % (N, name, ms, B * N * d * 4 / ms / 1e6, B * N * d * 4 / ms / 1e6 / 65.533, err), flush=True)
lib.ur_score_loss_set_bulk(9)
