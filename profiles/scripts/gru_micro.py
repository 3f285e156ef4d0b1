"""Micro-benchmark of the persistent GRU recurrence kernels at the BASELINE c3 shape (B=2048, L=100, h=256)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from unirec_b200 import ops
B, L, H = int(os.environ.get('B', 2048)), int(os.environ.get('L', 100)), int(os.environ.get('H', 256))
dev = 'cuda'
torch.manual_seed(0)
k = 1.0 / H ** 0.5
w_hh = (torch.rand(3 * H, H, device=dev) * 2 - 1) * k
b_hh = (torch.rand(3 * H, device=dev) * 2 - 1) * k
gi = torch.randn(B, L, 3 * H, device=dev) * 0.5
hs = torch.zeros(L + 1, B, H, device=dev)
save = torch.empty(L, B, 4 * H, device=dev)
whh_t = ops.transpose(w_hh, torch.empty(H, 3 * H, device=dev))
dgi = torch.empty(B, L, 3 * H, device=dev)
dgh = torch.empty(L, B, 3 * H, device=dev)
dh = torch.randn(B, H, device=dev)
for it in range(3):
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    ops.gru_seq_fwd(gi, whh_t, b_hh, hs, save, B, L, H)
    b.record()
    ops.gru_seq_bwd(dh, save, hs, w_hh, dgi, dgh, B, L, H)
    c.record()
    torch.cuda.synchronize()
    print('fwd %.3f ms  bwd %.3f ms' % (a.elapsed_time(b), b.elapsed_time(c)))
