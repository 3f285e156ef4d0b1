"""Owner-side partial-softmax kernel (ur_score_partial_f32) on ONE GPU acting as rank 0 of 8: variants via UR_PARTIAL_VARIANT."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from unirec_b200 import ops
W = 8
for d, N, S in ((128, 1025, 8192), (256, 4097, 4096)):
    V = 10_000_000
    table = torch.randn(V // W, d, device='cuda') * 0.02
    g = torch.Generator(device='cuda').manual_seed(1)
    ids = [(torch.randint(1, V, (S, N), device='cuda', generator=g).to(torch.int32)) for _ in range(3)]
    for t in ids:
        t[:, 0] |= -(1 << 31)
    u = torch.randn(S, d, device='cuda') * 0.02
    z = torch.empty(S, N, device='cuda'); state = torch.empty(S, 4 + 2 * d, device='cuda')
    def run(i):
        ops.score_partial(table, u, ids[i % 3], W, 0, z, state)
    for i in range(3): run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): run(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    nbytes = S * N / W * d * 4
    print('variant %s d=%d N=%d S=%d: %.4f ms  %.0f GB/s of owned rows (%.1f%% of 6539)  checksum %.6f' % (os.environ.get('UR_PARTIAL_VARIANT', '0'), d, N, S, ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / 65.395, float(state[:, 1].sum())), flush=True)
    del table
